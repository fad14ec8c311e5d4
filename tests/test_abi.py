"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports
every symbol include/soket_b200.h declares, fails loudly (no CPU fallback) when asked to
compute without a device; the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "soket_b200.h")
LIB = os.path.join(ROOT, "soket_b200", "lib", "libsoketb200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    assert os.path.exists(LIB), "libsoketb200.so is not built: python soket_b200/build.py"
    return ctypes.CDLL(LIB)


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("sk_init", "sk_malloc", "sk_ewise_binary", "sk_copy", "sk_reduce", "sk_argreduce",
                 "sk_gather_rows", "sk_matmul", "sk_linear_fwd", "sk_layernorm_fwd", "sk_layernorm_bwd",
                 "sk_batchnorm_fwd", "sk_batchnorm_bwd", "sk_softmax_ce_fwd_bwd", "sk_sgd_step",
                 "sk_adam_step", "sk_nccl_allreduce"):
        assert must in names
    assert len(names) >= 60


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    lib.sk_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sk_version()
    n = ctypes.c_int(-1)
    assert lib.sk_device_count(ctypes.byref(n)) == 0
    if n.value > 0:
        pytest.skip("a GPU is visible; the fail-loud path is exercised on the CPU box")
    lib.sk_last_error.restype = ctypes.c_char_p
    assert lib.sk_init(0) != 0
    assert b"no CUDA device" in lib.sk_last_error()
    p = ctypes.c_void_p()
    assert lib.sk_malloc(ctypes.c_size_t(64), ctypes.byref(p)) != 0
    import soket_b200 as sk
    with pytest.raises(RuntimeError):
        sk.zeros((4,), "float32")
    with pytest.raises(RuntimeError):
        sk.add(1.0, 2.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "soket_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".pyx", ".pxd", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/_ref" in src and f != "compat.py":
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_compat_module_covers_the_seam():
    """Every name the reference takes from its GPU array library
    (soket/tensor/ops/intern.pyx:48-76, soket/backend/device.pyx:64-71,
    soket/tensor/tensor.pyx:225-226) exists on the compat module."""
    import soket_b200.compat as compat
    mod = compat.make_module("cupy")
    intern = ["array", "add", "negative", "subtract", "multiply", "divide", "power", "sum", "mean", "max",
              "min", "argmax", "argmin", "reshape", "broadcast_to", "log", "exp", "matmul", "copy", "equal",
              "not_equal", "greater", "greater_equal", "less", "less_equal", "transpose", "maximum",
              "squeeze", "stack"]
    assert len(intern) == 29
    for n in intern + ["zeros", "ones", "eye", "empty", "full", "ndarray", "asnumpy"]:
        assert callable(getattr(mod, n)) or isinstance(getattr(mod, n), type), n
    for n in ("uniform", "normal", "binomial"):
        assert callable(getattr(mod.random, n))
    d = mod.cuda.Device(0)
    assert d == mod.cuda.Device(0) and d != mod.cuda.Device(1)
    for n in ("use", "synchronize", "__enter__", "__exit__"):
        assert hasattr(d, n)
