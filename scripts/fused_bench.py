"""Fused-kernel microbench at the benchmark's shapes (B = 8192, H = 4096):
LayerNorm forward / backward in the variants the wide MLPResNet step launches, the operand-split
passes, Adam, the fused elementwise interpreter -- achieved ALGORITHMIC HBM GB/s per launch (CUDA
events, L2 flushed between launches) against MEASURED_PEAKS.json.

    python scripts/fused_bench.py [--rows 8192] [--cols 4096] [--reps 10] [--only ln] > profiles/r2_fused_bench.md
    ncu --set full ... python scripts/fused_bench.py --reps 1 --only ln      # one launch of each for a capture
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soket_b200 as sk  # noqa: E402
from soket_b200 import _fused as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--cols", type=int, default=4096)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
ap.add_argument("--json", default="")
args = ap.parse_args()
sk.init(0)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]
R, C = args.rows, args.cols
n = R * C
rng = np.random.default_rng(0)
x = sk.array((rng.standard_normal((R, C)) * 2 + 0.5).astype("float32"))
res = sk.array(rng.standard_normal((R, C)).astype("float32"))
adj = sk.array(rng.standard_normal((R, C)).astype("float32"))
g = sk.array((rng.random(C) + 0.5).astype("float32"))
b = sk.array((rng.standard_normal(C) * 0.1).astype("float32"))
w = sk.array((rng.standard_normal((C, C)) * 0.02).astype("float32"))
gw = sk.array((rng.standard_normal((C, C)) * 1e-3).astype("float32"))
res_split = sk.split_f16(res)
_, mean, rstd = F.layernorm_fwd(x, g, b, None, 1e-5, True)
y_res, mean2, rstd2 = F.layernorm_fwd(x, g, b, res, 1e-5, True)
sk.random.seed(1)
_, mean3, rstd3, seed3 = F.layernorm_dropout_fwd(x, g, b, 1e-5, True, 0.99)
m_state, v_state = sk.zeros((C, C), "float32"), sk.zeros((C, C), "float32")

CASES = {
    # name: (callable, algorithmic bytes)
    "ln_fwd plain": (lambda: F.layernorm_fwd(x, g, b, None, 1e-5, False), 8 * n),
    "ln_fwd relu+dropout (block LN1)": (lambda: F.layernorm_dropout_fwd(x, g, b, 1e-5, True, 0.99), 8 * n),
    "ln_fwd relu+dropout +split": (lambda: F.layernorm_dropout_fwd(x, g, b, 1e-5, True, 0.99, True), 12 * n),
    "ln_fwd residual+relu (block LN2)": (lambda: F.layernorm_fwd(x, g, b, res, 1e-5, True), 12 * n),
    "ln_fwd residual+relu +split": (lambda: F.layernorm_fwd(x, g, b, res, 1e-5, True, True, res_split), 16 * n),
    "ln_bwd relu+dropout (LN1)": (lambda: F.layernorm_dropout_bwd(adj, x, g, b, mean3, rstd3, True, 0.99, 1 / 0.99, seed3), 12 * n),
    "ln_bwd relu+dropout +absmax": (lambda: F.layernorm_dropout_bwd(adj, x, g, b, mean3, rstd3, True, 0.99, 1 / 0.99, seed3, True, None, None, True), 12 * n),
    "ln_bwd residual mask + dresidual (LN2)": (lambda: F.layernorm_bwd(adj, x, g, b, mean2, rstd2, y_res, 2, True), 20 * n),
    "ln_bwd plain": (lambda: F.layernorm_bwd(adj, x, g, b, mean, rstd), 12 * n),
    "split_f16 (absmax pass + split)": (lambda: sk.split_f16(x), 12 * n),
    "split_f16 + column sums (adjoint)": (lambda: sk.split_f16(adj, True), 12 * n),
    "split_f16 weight (4096 x 4096)": (lambda: sk.split_f16(w), 12 * C * C),
    "adam step (4096 x 4096)": (lambda: F.adam_step([w], [gw], [m_state], [v_state], 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, False), 28 * C * C),
    "accumulate_": (lambda: F.accumulate_(res, adj), 12 * n),
}


def gpu_time(fn, reps):
    for _ in range(2 if reps > 1 else 0):
        fn()
    ts = []
    for _ in range(reps):
        sk.flush_l2()
        e0, e1 = sk.Event(), sk.Event()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_ms(e1))
    return float(np.median(ts))


print(f"# fused kernels at ({R}, {C}) fp32, B200; measured HBM copy peak {PEAK:.0f} GB/s\n")
print("| kernel | ms | algorithmic bytes | GB/s | frac of measured peak | frac of 8 TB/s |")
print("|---|---:|---:|---:|---:|---:|")
out = []
for name, (fn, nbytes) in CASES.items():
    if args.only and args.only not in name:
        continue
    ms = gpu_time(fn, args.reps)
    gbs = nbytes / ms / 1e6
    out.append({"kernel": name, "ms": ms, "bytes": nbytes, "gbs": gbs})
    print(f"| {name} | {ms:.4f} | {nbytes / 1e6:.0f} MB | {gbs:.0f} | {gbs / PEAK:.2f} | {gbs / 8000:.2f} |")
if args.json:
    json.dump({"peak_gbs": PEAK, "rows": R, "cols": C, "results": out}, open(args.json, "w"))
