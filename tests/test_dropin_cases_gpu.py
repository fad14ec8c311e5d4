"""The drop-in claim beyond MLPResNet training: every Tensor-op and nn-module case, run by the
UNMODIFIED reference's own Tensor / autodiff / nn code on `soket.gpu()` with soket_b200 in the
CuPy seam, reproduces what the same reference computed on its CPU device (the committed goldens):
1e-5 relative on fp32 values, exact shapes / dtypes / integer and bool results."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_on_the_backend_reproduces_its_cpu_results(sk, ref_soket):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dropin_cases.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"rc {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    results = json.loads(r.stdout.strip().splitlines()[-1])
    assert "__fatal__" not in results, results
    assert len(results) >= 90
    bad = {k: v for k, v in results.items() if v != "ok"}
    assert not bad, json.dumps(bad, indent=1)
