"""soket_b200.transforms -- soket/transforms.pyx:7-89 for the GPU device.

``Transform`` is the callable base class (``transform(*args)`` overridden by
subclasses, at least one positional argument required, transforms.pyx:27-37);
``ToTensor`` turns one NumPy sample into a device ``Tensor``
(transforms.pyx:73-87: anything that is not an ``ndarray`` goes through
``numpy.array`` first, so a ``numpy.uint8`` label becomes a 0-d uint8 tensor).
"""
import numpy as _np

from soket_b200.engine import Tensor


class Transform:
    def transform(self, *args):
        raise NotImplementedError()

    def __call__(self, *args):
        if len(args) == 0:
            raise ValueError('Expected atleast one positional argument!')
        return self.transform(*args)


class ToTensor(Transform):
    def transform(self, X, *rest):
        if type(X) is not _np.ndarray:
            X = _np.array(X)
        return Tensor(X)
