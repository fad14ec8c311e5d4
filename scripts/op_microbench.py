"""Backend op microbench (BASELINE.json configs[1]): fp32 tensors from
2^20 to 2^30 elements, achieved ALGORITHMIC HBM GB/s per op (CUDA events around each launch,
L2 flushed between iterations), next to the same NumPy call on the host cores.

    python scripts/op_microbench.py [--max-log2 30] [--cpu-max-log2 24] > gpurun_out/op_microbench.md
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import soket_b200 as sk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--max-log2", type=int, default=30)
ap.add_argument("--min-log2", type=int, default=20)
ap.add_argument("--cpu-max-log2", type=int, default=24)
ap.add_argument("--json", default="gpurun_out/op_microbench.json")
args = ap.parse_args()
sk.init(0)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
PEAK = peaks["hbm_gbs"]
C = 4096


def gpu_time(fn, reps):
    for _ in range(3):
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


def cpu_time(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def cases(n, xp, a, b):
    """name -> (callable, algorithmic bytes).  xp is soket_b200 (device) or numpy (oracle calls)."""
    r = n // C
    a2 = xp.reshape(a, (r, C))
    col = xp.reshape(a, (n,))[:r].reshape(r, 1)
    row = xp.reshape(a, (n,))[:C]
    red_sum = (lambda x, ax, kd=False: xp.sum(x, ax, "float32", None, kd))
    red_max = (lambda x, ax, kd=False: xp.max(x, ax, None, kd))
    return {
        "a+b": (lambda: xp.add(a, b), 12 * n),
        "a*b": (lambda: xp.multiply(a, b), 12 * n),
        "a+s": (lambda: xp.add(a, 1.5), 8 * n),
        "a*s": (lambda: xp.multiply(a, 1.5), 8 * n),
        "maximum(a,0)": (lambda: xp.maximum(a, 0), 8 * n),
        "exp(a)": (lambda: xp.exp(a), 8 * n),
        "a2d+row (bias)": (lambda: xp.add(a2, row), 8 * n + 4 * C),
        "broadcast col->(r,4096)": (lambda: xp.ascontiguousarray(xp.broadcast_to(col, (r, C))), 4 * n + 4 * r),
        "broadcast row->(r,4096)": (lambda: xp.ascontiguousarray(xp.broadcast_to(row, (r, C))), 4 * n + 4 * C),
        "compact a2d.T": (lambda: xp.ascontiguousarray(a2.T), 8 * n),
        "reduce_sum full": (lambda: red_sum(a, None), 4 * n),
        "reduce_sum last axis": (lambda: red_sum(a2, (1,)), 4 * n + 4 * r),
        "reduce_sum first axis": (lambda: red_sum(a2, (0,)), 4 * n + 4 * C),
        "reduce_max full": (lambda: red_max(a, None), 4 * n),
        "reduce_max last axis": (lambda: red_max(a2, (1,)), 4 * n + 4 * r),
        "reduce_max first axis": (lambda: red_max(a2, (0,)), 4 * n + 4 * C),
    }


results = []
sizes = list(range(args.min_log2, args.max_log2 + 1, 2))
print(f"# op microbench, fp32, B200 (measured HBM copy peak {PEAK:.0f} GB/s; north-star peak 8000 GB/s); host cores: {os.cpu_count()}\n")
print("| op | " + " | ".join(f"2^{k}" for k in sizes) + " | best frac of measured | best frac of 8 TB/s | CPU GB/s (2^%d) |" % min(args.cpu_max_log2, sizes[-1]))
print("|---|" + "---:|" * (len(sizes) + 3))
table = {}
cpu_col = {}
for k in sizes:
    n = 1 << k
    a = sk.random.uniform(0, 1, (n,), dtype="float32")
    b = sk.random.uniform(0, 1, (n,), dtype="float32")
    reps = 10 if k <= 26 else 5
    for name, (fn, nbytes) in cases(n, sk, a, b).items():
        try:
            ms = gpu_time(fn, reps)
            gbs = nbytes / ms / 1e6
        except Exception as e:   # an op that cannot run at this size is reported, not hidden
            print(f"<!-- {name} 2^{k}: {e} -->")
            ms, gbs = float("nan"), float("nan")
        table.setdefault(name, {})[k] = gbs
        results.append({"op": name, "log2_n": k, "ms": ms, "gbs": gbs, "bytes": nbytes})
    del a, b
    sk.empty_cache()
    if k == min(args.cpu_max_log2, sizes[-1]):
        rng = np.random.default_rng(0)
        ha, hb = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
        for name, (fn, nbytes) in cases(n, np, ha, hb).items():
            cpu_col[name] = nbytes / cpu_time(fn) / 1e6
for name, row in table.items():
    best = np.nanmax(list(row.values()))
    print(f"| {name} | " + " | ".join(f"{row[k]:.0f}" for k in sizes) +
          f" | {best / PEAK:.2f} | {best / 8000:.2f} | {cpu_col.get(name, float('nan')):.1f} |")
os.makedirs(os.path.dirname(args.json), exist_ok=True)
json.dump({"peak_gbs": PEAK, "results": results, "cpu_gbs": cpu_col, "cpu_cores": os.cpu_count()}, open(args.json, "w"))
