"""Run the Tensor-op and nn-module case tables (tests/tensor_op_cases.py,
tests/nn_module_cases.py) through the UNMODIFIED reference on `soket.gpu()` -- its own Tensor /
autodiff / nn code, with soket_b200 in the seam where CuPy sits (soket_b200.compat) -- and
compare with the goldens the same reference produced on its CPU device.  Prints one JSON object
{case: "ok" | "<what differs>"}; run in its own interpreter by tests/test_dropin_cases_gpu.py
(a crash inside the reference then fails one test instead of the session)."""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_model                       # noqa: E402

soket = ref_model.import_reference()
if soket is None:
    print(json.dumps({"__fatal__": "oracle/_ref is not built"}))
    sys.exit(0)
import soket.nn as nn                               # noqa: E402
from nn_module_cases import CASES as NN_CASES       # noqa: E402
from tensor_op_cases import CASES, INT_CASES, make_inputs   # noqa: E402

GOLD_OPS = np.load(os.path.join(ROOT, "tests", "golden", "tensor_ops.npz"))
GOLD_NN = np.load(os.path.join(ROOT, "tests", "golden", "nn_modules.npz"))
SKIP_OPS = {c["case"] for c in json.loads(str(GOLD_OPS["__crashes__"]))}
SKIP_NN = {c["case"] for c in json.loads(str(GOLD_NN["__crashes__"]))}
cpu, gpu = soket.cpu(), soket.gpu()


def to_numpy(t):
    t = soket.Tensor(t, cpu)
    if len(t.shape) == 0:
        return np.array(t.item(), dtype=str(t.dtype))
    buf = np.zeros(t.shape, dtype=str(t.dtype))
    view = soket.Tensor.from_numpy(buf)
    view[tuple(slice(None) for _ in t.shape)] = t
    return buf


def differs(got, want, what):
    if got.shape != want.shape:
        return f"{what}: shape {got.shape} vs {want.shape}"
    if got.dtype != want.dtype:
        return f"{what}: dtype {got.dtype} vs {want.dtype}"
    if want.dtype.kind in "biu":
        return None if np.array_equal(got, want) else f"{what}: values differ"
    if want.size == 0:
        return None
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max())
    return None if err <= 1e-5 * scale + 1e-7 else f"{what}: err {err:.3e} at scale {scale:.3e}"


def seed_of(name):
    return zlib.crc32(name.encode())


results = {}
with gpu:
    for idx, (name, shapes, fn) in enumerate(CASES + INT_CASES):
        if name in SKIP_OPS:
            continue
        try:
            xs = [soket.Tensor(a, requires_grad=True) for a in make_inputs(shapes, seed_of(name))]
            assert all("GPU" in str(x.device) for x in xs)
            out = fn(soket, *xs)
            bad = differs(to_numpy(out), GOLD_OPS[f"{name}/out"], "out")
            if bad is None and idx < len(CASES) and out.requires_grad and f"{name}/backward_error" not in GOLD_OPS.files:
                w = np.random.default_rng(seed_of(name) ^ 0x5EED).standard_normal(tuple(out.shape)).astype("float32")
                (out * soket.Tensor(w)).sum().backward()
                for i, x in enumerate(xs):
                    key = f"{name}/grad{i}"
                    if key in GOLD_OPS.files and bad is None:
                        bad = "missing grad" if x.grad is None else differs(to_numpy(x.grad), GOLD_OPS[key], f"grad{i}")
            results["op:" + name] = bad or "ok"
        except Exception as e:
            results["op:" + name] = f"raised {type(e).__name__}: {e}"[:300]
    for fn in NN_CASES:
        name = fn.__name__
        if name in SKIP_NN:
            continue
        try:
            res = fn(soket, nn, np.random.default_rng(seed_of(name)))
            bad = None
            for k, v in res.items():
                want = GOLD_NN[f"{name}/{k}"]
                if isinstance(v, soket.Tensor):
                    bad = bad or differs(to_numpy(v), want, k)
                elif want.dtype.kind in "US":
                    if k == "str" or k.endswith("_str"):
                        continue            # printed with the GPU device suffix here
                    if str(v) != str(want):
                        bad = bad or f"{k}: {v!r} vs {str(want)!r}"
                elif v != want.item():
                    bad = bad or f"{k}: {v!r} vs {want.item()!r}"
            results["nn:" + name] = bad or "ok"
        except Exception as e:
            results["nn:" + name] = f"raised {type(e).__name__}: {e}"[:300]
print(json.dumps(results))
