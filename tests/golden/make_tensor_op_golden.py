"""Generate tests/golden/tensor_ops.npz by running the BUILT reference (oracle/_ref, CPU /
NumPy device) on every case of tests/tensor_op_cases.py: forward value and, for
differentiable cases, the input gradients of sum(out * w) with a fixed random w.

Each case runs in its own interpreter because the reference itself crashes on some inputs
(e.g. logsumexp(..., keepdims=True) dereferences a freed value cache): such cases are
recorded under `crashes` and the GPU test only checks that soket_b200 handles them.

    python tests/golden/make_tensor_op_golden.py            # needs /root/reference -> oracle/_ref
"""
import json
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def seed_of(name):
    return zlib.crc32(name.encode())


def weights_for(shape, name):
    return np.random.default_rng(seed_of(name) ^ 0x5EED).standard_normal(shape).astype("float32")


def run_case(index, out_path):
    from oracle import ref_model
    from tensor_op_cases import CASES, INT_CASES, MULTI_CASES, make_inputs
    soket = ref_model.import_reference()

    def to_numpy(t):
        if len(t.shape) == 0:
            return np.array(t.item(), dtype=str(t.dtype))
        buf = np.zeros(t.shape, dtype=str(t.dtype))
        view = soket.Tensor.from_numpy(buf)
        view[tuple(slice(None) for _ in t.shape)] = t
        return buf
    cases = CASES + INT_CASES
    if index >= len(cases):               # dtype / creation tables: a list of results per case
        name, fn = MULTI_CASES[index - len(cases)]
        res = {}
        for i, o in enumerate(fn(soket, None)):
            if isinstance(o, str):
                res[f"r{i:02d}"] = np.array(o)
            else:
                res[f"r{i:02d}"] = to_numpy(o)
                res[f"r{i:02d}_dtype"] = np.array(str(o.dtype))      # the Tensor's dtype TAG
        np.savez(out_path, **res)
        return
    name, shapes, fn = cases[index]
    xs = [soket.Tensor(a, requires_grad=True) for a in make_inputs(shapes, seed_of(name))]
    out = fn(soket, *xs)
    res = {"out": to_numpy(out)}
    if index < len(CASES) and out.requires_grad:
        np.savez(out_path, **res)            # the forward value survives a crash in backward
        w = soket.Tensor(weights_for(res["out"].shape, name))
        try:
            (out * w).sum().backward()
        except Exception as e:               # e.g. batched matmul: `.T` reverses ALL axes (quirk Q7)
            res["backward_error"] = np.array(f"{type(e).__name__}: {e}"[:300])
        else:
            for i, x in enumerate(xs):
                if x.grad is not None:
                    res[f"grad{i}"] = to_numpy(x.grad)
    np.savez(out_path, **res)


def main():
    from tensor_op_cases import CASES, INT_CASES, MULTI_CASES
    cases = CASES + INT_CASES + [(n, None, f) for n, f in MULTI_CASES]
    golden, crashes = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        for i, (name, _, _) in enumerate(cases):
            path = os.path.join(tmp, f"{i}.npz")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i), path],
                               capture_output=True, text=True)
            if r.returncode != 0:
                last = r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""
                crashes.append({"case": name, "returncode": r.returncode, "stderr_tail": last[:200],
                                "forward_saved": os.path.exists(path)})
                print(f"  {name}: reference failed (rc {r.returncode}) {last[:120]}")
                if not os.path.exists(path):
                    continue
            with np.load(path) as z:
                for k in z.files:
                    golden[f"{name}/{k}"] = z[k]
    golden["__crashes__"] = np.array(json.dumps(crashes))
    np.savez_compressed(os.path.join(HERE, "tensor_ops.npz"), **golden)
    print(f"wrote tensor_ops.npz: {len(cases) - len(crashes)} cases, {len(crashes)} reference failures")


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--case":
        run_case(int(sys.argv[2]), sys.argv[3])
    else:
        main()
