"""soket_b200.compat -- plug the backend into the seam where CuPy sits.

The reference selects its GPU array library by name: ``import cupy`` /
``from cupy.cuda import Device`` in soket/backend/device.pyx:13-20, ``import cupy
as cp`` in soket/tensor/ops/intern.pyx:45-47 and soket/tensor/tensor.pyx:221-226.
``install()`` registers module objects under those names whose attributes are
this package's functions, so an UNMODIFIED Soket build dispatches every
``soket.gpu()`` tensor operation to the sm_100a kernels:

    import soket_b200.compat as compat
    compat.install()          # before `import soket`
    import soket
    with soket.gpu():
        ...

(INTEGRATION.md shows the equivalent two-line source change for a maintainer
who prefers an explicit import.)  Nothing here touches CuPy itself; if a real
CuPy is installed, ``install(force=True)`` shadows it for this process.
"""
from __future__ import annotations

import sys
import types

import soket_b200 as _b
from soket_b200 import _core

# the 29 intern-table slots (soket/tensor/ops/intern.pyx:48-76) + Device creation
# functions (soket/backend/device.pyx:64-71) + the names tensor.pyx uses
_SURFACE = [
    "array", "add", "negative", "subtract", "multiply", "divide", "power", "sum", "mean",
    "max", "min", "argmax", "argmin", "reshape", "broadcast_to", "log", "exp", "matmul",
    "copy", "equal", "not_equal", "greater", "greater_equal", "less", "less_equal",
    "transpose", "maximum", "squeeze", "stack",
    "zeros", "ones", "eye", "empty", "full", "ndarray", "asnumpy", "asarray", "random",
    "minimum", "sqrt", "absolute", "zeros_like", "ones_like", "empty_like", "expand_dims",
    "ascontiguousarray",
]


def make_module(name: str = "cupy") -> types.ModuleType:
    mod = types.ModuleType(name, "soket_b200 presented under the array-library name Soket imports")
    for n in _SURFACE:
        setattr(mod, n, getattr(_b, n))
    cuda = types.ModuleType(name + ".cuda")
    cuda.Device = _core._DeviceHandle
    cuda.is_available = _b.is_available
    cuda.runtime = types.SimpleNamespace(getDeviceCount=_b.device_count,
                                         deviceSynchronize=_b.synchronize)
    mod.cuda = cuda
    mod.__version__ = "soket_b200-" + _b.__version__
    mod.__soket_b200__ = True
    return mod


def install(name: str = "cupy", force: bool = False) -> types.ModuleType:
    """Register soket_b200 as the module Soket imports for its GPU backend."""
    existing = sys.modules.get(name)
    if existing is not None and not getattr(existing, "__soket_b200__", False) and not force:
        raise RuntimeError(f"a different '{name}' module is already imported; pass force=True to shadow it")
    mod = make_module(name)
    sys.modules[name] = mod
    sys.modules[name + ".cuda"] = mod.cuda
    return mod
