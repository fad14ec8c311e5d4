import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import soket_b200
        return soket_b200.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def sk():
    """The device backend, initialised on cuda:0.  GPU tests FAIL (not skip) if the
    extension is missing: there is no CPU fallback to fall back to."""
    import soket_b200
    if soket_b200.device_count() == 0:
        pytest.skip("no CUDA device visible")
    soket_b200.init(0)
    return soket_b200


@pytest.fixture(scope="session")
def ref_soket():
    """The built reference (oracle/_ref) or skip."""
    from oracle import ref_model
    mod = ref_model.import_reference()
    if mod is None:
        pytest.skip("oracle/_ref not built (run python oracle/build_ref.py where /root/reference exists)")
    return mod
