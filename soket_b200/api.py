"""soket_b200.api -- the names a Soket user imports, on the sm_100a backend.

    import soket_b200.api as soket
    import soket_b200.api.nn as nn            # (or: from soket_b200 import nn)
    from soket_b200.api.optim import Adam

mirrors ``import soket`` / ``soket.nn`` / ``soket.optim`` of the reference
(soket/__init__.py:1-7) for code that runs entirely on the GPU device.
"""
from soket_b200.engine import (  # noqa: F401
    Tensor, DType, Device, DeviceType, gpu, cpu, promote_types, lazy, LazyState, lazy_stats,
    float16, float32, float64, int8, uint8, int16, uint16, int32, uint32, int64, uint64, bool_,
    rand, randn, randb, zeros, ones, empty, full, one_hot,
    zeros_like, ones_like, one_like, empty_like, rand_like, randn_like,
    log, exp, logsumexp, stack, set_leaf_grad_hook,
)
from soket_b200 import nn, optim, transforms, utils  # noqa: F401
import soket_b200.utils.data  # noqa: F401,E402  (soket.utils.data: Dataset, DataLoader, MNIST)

bool = bool_  # soket exports the dtype under the name `bool` (soket/dtype.pyx:137-151)
