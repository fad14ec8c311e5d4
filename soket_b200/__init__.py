"""soket_b200 -- a from-scratch B200 (sm_100a) device backend for Soket.

The package exposes, at module level, exactly the array-library surface the
reference reaches through its drop-in seam: the
``Device._backend`` functions (soket/backend/device.pyx:52-71), the 29 callables
of the intern table's GPU column (soket/tensor/ops/intern.pyx:45-76) and the
``ndarray`` protocol.  ``soket_b200.compat.install()`` registers it under the
name the reference imports (``cupy``), which makes the UNMODIFIED reference run
on these kernels; ``soket_b200.engine`` is the resident / fused training path
(fused Linear+bias+ReLU, LayerNorm, BatchNorm, softmax-CE, multi-tensor
SGD/Adam, NCCL data parallel).

There is no CPU fallback and no PyTorch / CuPy / Triton on the device path:
if the CUDA extension is missing the import fails, and every call needs a
CUDA device.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
if not _os.path.exists(_os.path.join(_here, "lib", "libsoketb200.so")):
    raise ImportError(
        "soket_b200: lib/libsoketb200.so is not built -- run `python soket_b200/build.py` "
        "(nvcc, sm_100a).  There is no CPU fallback.")

try:
    from . import _core
except ImportError as _e:  # pragma: no cover - build problem, fail loudly
    raise ImportError(
        f"soket_b200: the Cython extension soket_b200._core failed to import ({_e}); "
        "run `python soket_b200/build.py`.  There is no CPU fallback.") from _e

from ._core import (  # noqa: F401,E402
    ndarray, array, asarray, asnumpy, copy, to_bf16, bfloat16,
    empty, zeros, ones, full, eye, zeros_like, ones_like, empty_like,
    transpose, broadcast_to, reshape, squeeze, expand_dims, ascontiguousarray, stack,
    add, subtract, multiply, divide, true_divide, power, maximum, minimum, negative,
    exp, log, sqrt, absolute, abs,
    equal, not_equal, greater, greater_equal, less, less_equal,
    sum, mean, max, min, amax, amin, argmax, argmin,
    matmul, linear, linear_bwd, relu_backward, one_hot, set_matmul_algo,
    MM_AUTO, MM_SIMT, MM_TF32X3, MM_TF32, MM_BF16, MM_F16X3,
    random, device_count, is_available, init, synchronize, launch_count, flush_l2,
    memory_stats, empty_cache, version, Event, Graph, PinnedBuffer,
    arena_create, arena_begin, arena_end, arena_destroy, rng_epoch_advance, is_capturing, prefetch_wait,
    profile_enable, profile_reset, profile_collect,
    SplitMat, split_f16, get_split, gemm_split, gemm_split_supported,
    launch_stream, STREAM_COMPUTE, STREAM_COMM, STREAM_COPY, STREAM_OPT,
)

__version__ = "0.1.0"

LIB_PATH = _os.path.join(_here, "lib", "libsoketb200.so")
