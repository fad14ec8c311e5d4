"""Generate tests/golden/nn_modules.npz by running the BUILT reference (oracle/_ref, CPU) on
every case of tests/nn_module_cases.py, one interpreter per case (the reference can crash).

    python tests/golden/make_nn_module_golden.py
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


def run_case(index, out_path):
    from oracle import ref_model
    from nn_module_cases import CASES
    soket = ref_model.import_reference()
    import soket.nn as nn

    def to_numpy(t):
        if len(t.shape) == 0:
            return np.array(t.item(), dtype=str(t.dtype))
        buf = np.zeros(t.shape, dtype=str(t.dtype))
        view = soket.Tensor.from_numpy(buf)
        view[tuple(slice(None) for _ in t.shape)] = t
        return buf
    fn = CASES[index]
    res = fn(soket, nn, np.random.default_rng(zlib.crc32(fn.__name__.encode())))
    out = {}
    for k, v in res.items():
        if isinstance(v, soket.Tensor):
            out[k] = to_numpy(v)
        elif v is None:
            out[k] = np.array("None")
        else:
            out[k] = np.array(v)
    np.savez(out_path, **out)


def main():
    from nn_module_cases import CASES
    golden, crashes = {}, []
    with tempfile.TemporaryDirectory() as tmp:
        for i, fn in enumerate(CASES):
            path = os.path.join(tmp, f"{i}.npz")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i), path],
                               capture_output=True, text=True)
            if r.returncode != 0:
                last = r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""
                crashes.append({"case": fn.__name__, "returncode": r.returncode, "stderr_tail": last[:300]})
                print(f"  {fn.__name__}: reference failed (rc {r.returncode}) {last[:200]}")
                continue
            with np.load(path) as z:
                for k in z.files:
                    golden[f"{fn.__name__}/{k}"] = z[k]
    golden["__crashes__"] = np.array(json.dumps(crashes))
    np.savez_compressed(os.path.join(HERE, "nn_modules.npz"), **golden)
    print(f"wrote nn_modules.npz: {len(CASES) - len(crashes)} cases, {len(crashes)} reference failures")


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--case":
        run_case(int(sys.argv[2]), sys.argv[3])
    else:
        main()
