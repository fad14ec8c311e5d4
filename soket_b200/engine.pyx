# cython: language_level=3, boundscheck=True, wraparound=True, cdivision=True
"""soket_b200.engine -- Soket's Tensor / autodiff API on resident device arrays.

Host-side mirror of the reference's L3/L4 layers for the GPU device
(soket/tensor/tensor.pyx, soket/autodiff.pyx, soket/tensor/creation.pyx,
soket/tensor/detached.pyx), freshly written over ``soket_b200._core``:

  * same public names and argument meaning (``Tensor``, ``rand/randn/randb/zeros/
    ones/empty/full/one_hot/*_like``, ``log/exp/logsumexp``, ``stack``), same
    shape / dtype inference and error behaviour, same reference quirks where they
    change results (Q5 ``mean`` backward scale, Q6 ``<=``, Q7 ``.T`` reverses all
    axes: SURVEY.md section 7);
  * tensors stay resident: no host bounce anywhere on the path (tensor.pyx:384-442
    bounces through ``asnumpy`` only for cross-device moves, as here);
  * reverse-mode AD with the reference's algorithm (topological order, partial
    adjoints, broadcast-compat reduction: autodiff.pyx:30-164) but partial
    adjoints are accumulated IN PLACE on device, add/sub backward ALIAS the
    adjoint instead of copying it, and a leaf's gradient is finalised the moment
    its last partial arrives (the hook data-parallel all-reduce hangs on).

The fused ops used by ``soket_b200.nn`` (linear+bias+relu, layer/batch norm
(+relu)(+residual), softmax-CE, dropout) are defined at the bottom.
"""
from libc.stdint cimport int64_t
from cpython.ref cimport PyObject

import numpy as np

from soket_b200._abi cimport *
from soket_b200._core cimport ndarray, Buffer, SplitMat, _new_array, _check, _fptr, _as_device
from soket_b200 import _core as B
from soket_b200 import _fused as F
import os as _os_env


# ============================================================================ dtypes
class DType:
    """soket/dtype.pyx:17-113: named dtype with promotion index."""
    _names = ['float16', 'float32', 'float64', 'int8', 'uint8', 'int16', 'uint16', 'int32',
              'uint32', 'int64', 'uint64', 'bool']

    def __init__(self, name):
        if isinstance(name, DType):
            name = name.name
        name = str(name)
        if name not in DType._names:
            raise ValueError(f"Unsupported datatype '{name}'")
        self.name = name
        self._idx = DType._names.index(name)

    def __eq__(self, other):
        if isinstance(other, DType):
            return self._idx == other._idx
        if isinstance(other, str):
            return self.name == other
        return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def __str__(self):
        return self.name

    def __repr__(self):
        return f'soket.{self.name}'


float16 = DType('float16'); float32 = DType('float32'); float64 = DType('float64')
int8 = DType('int8'); uint8 = DType('uint8'); int16 = DType('int16'); uint16 = DType('uint16')
int32 = DType('int32'); uint32 = DType('uint32'); int64 = DType('int64'); uint64 = DType('uint64')
bool_ = DType('bool')
_DTYPES = {d.name: d for d in (float16, float32, float64, int8, uint8, int16, uint16, int32,
                               uint32, int64, uint64, bool_)}
_default_dtype = float32

# soket/dtype.pyx:173-185 (bool excluded; rows/cols in DType._names order)
_PROMO = [
    ['float16', 'float32', 'float64', 'float16', 'float16', 'float16', 'float16', 'float16', 'float16', 'float16', 'float16'],
    ['float32', 'float32', 'float64', 'float32', 'float32', 'float32', 'float32', 'float32', 'float32', 'float32', 'float32'],
    ['float64', 'float64', 'float64', 'float64', 'float64', 'float64', 'float64', 'float64', 'float64', 'float64', 'float64'],
    ['float16', 'float32', 'float64', 'int8', 'int16', 'int16', 'int32', 'int32', 'int64', 'int64', 'float32'],
    ['float16', 'float32', 'float64', 'int16', 'uint8', 'int16', 'uint16', 'int32', 'uint32', 'int64', 'uint64'],
    ['float16', 'float32', 'float64', 'int16', 'int16', 'int16', 'int32', 'int32', 'int64', 'int64', 'float32'],
    ['float16', 'float32', 'float64', 'int32', 'uint16', 'int32', 'uint16', 'int32', 'uint32', 'int64', 'uint64'],
    ['float16', 'float32', 'float64', 'int32', 'int32', 'int32', 'int32', 'int32', 'int64', 'int64', 'float32'],
    ['float16', 'float32', 'float64', 'int64', 'uint32', 'int64', 'uint32', 'int64', 'uint32', 'int64', 'uint64'],
    ['float16', 'float32', 'float64', 'int64', 'int64', 'int64', 'int64', 'int64', 'int64', 'int64', 'float32'],
    ['float16', 'float32', 'float64', 'float32', 'uint64', 'float32', 'uint64', 'float32', 'uint64', 'float32', 'uint64'],
]


def promote_types(a, b):
    """soket/dtype.pyx:189-204."""
    if a._idx == b._idx:
        return a
    if a._idx == bool_._idx:
        return b
    if b._idx == bool_._idx:
        return a
    return _DTYPES[_PROMO[a._idx][b._idx]]


def _scalar_dtype(s):
    """soket/dtype.pyx:157-167: int -> int32, float -> float32, bool -> bool."""
    if type(s) is int:
        return int32
    if type(s) is float:
        return float32
    if type(s) is bool:
        return bool_
    raise ValueError(f"Unsupported scalar '{s}'")


def _dt(x):
    """anything dtype-like -> DType (None -> default)."""
    if x is None:
        return _default_dtype
    if isinstance(x, DType):
        return x
    return _DTYPES[str(x)]


# ============================================================================ device
import enum as _enum


class DeviceType(_enum.IntEnum):
    """soket/backend/device.pxd:10-13."""
    CPU = 0
    GPU = 1


class Device:
    """soket/backend/device.pyx:33-249, GPU branch only: this package IS the GPU
    backend; CPU tensors belong to the reference's NumPy path."""

    def __init__(self, id=0):
        self._id = 0 if id is None else int(id)
        self._prev = None

    @property
    def type(self):
        return DeviceType.GPU       # device.pyx:166-170

    @property
    def id(self):
        return self._id

    def use(self):
        B.init(self._id)

    def sync(self):
        B.synchronize()

    def __enter__(self):
        self.use()
        return self

    def __exit__(self, *exc):
        return False

    def __eq__(self, other):
        return isinstance(other, Device) and other._id == self._id

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(('soket_b200.engine.Device', self._id))

    def __str__(self):
        return f'soket.Device(GPU, {self._id})'

    __repr__ = __str__


_gpu0 = Device(0)


def gpu(id=None):
    if id is None or id == 0:
        return _gpu0
    return Device(id)


def cpu():
    """device.pyx:270-272 returns the NumPy device; this package is the GPU device only."""
    raise RuntimeError('soket_b200 implements the GPU device only: CPU tensors belong to the reference\'s '
                       'NumPy backend (soket.cpu())')


# ---- lazy mode (tensor.pyx:24-51) -----------------------------------------------------------
# lazy() (tensor.pyx:24-51): the reference postpones Tensor._compute_data() until a value is needed
# (tensor.pyx:790-810); results are the same.  Here the switch opens a FUSION WINDOW: float32
# elementwise / scalar / unary nodes created while it is on are recorded, not launched, and the whole
# chain runs as ONE kernel (sk_ewise_fused) the moment some value is needed -- `.item()`, `.numpy()`,
# a reduction, a matmul, backward.  Everything else evaluates at construction as before.  See
# `_realize` below.
_LAZY_STATE = False


class LazyState:
    def __enter__(self):
        global _LAZY_STATE
        self._previous = _LAZY_STATE
        _LAZY_STATE = True

    def __exit__(self, *exc):
        global _LAZY_STATE
        _LAZY_STATE = self._previous
        return False


def lazy(enabled=None):
    global _LAZY_STATE
    if enabled is None:
        return LazyState()
    _LAZY_STATE = bool(enabled)


def lazy_enabled():
    return _LAZY_STATE


def _default_device():
    return _gpu0


# ============================================================================ shape helpers
cdef tuple _bshape(tuple a, tuple b, int nignore=0):
    """tensor.pyx:240-298 broadcast rule (RuntimeError on mismatch, like the reference)."""
    cdef tuple mx = a if len(a) >= len(b) else b
    cdef tuple mn = b if len(a) >= len(b) else a
    cdef int diff = len(mx) - len(mn), i
    cdef list out = list(mx)
    for i in range(len(mn) - nignore):
        mis = mn[i]; mas = mx[i + diff]
        if mis != 1 and mas != 1 and mis != mas:
            raise RuntimeError('Incompatible shapes!')
        out[i + diff] = mas if mas > mis else mis
    return tuple(out)


cdef tuple _norm_axes(object axes, int nd):
    """tensor.pxd:183-236: normalise, sort, validate; () means every axis."""
    if axes is None:
        return ()
    if not isinstance(axes, (tuple, list)):
        axes = (axes,)
    out = []
    for a in axes:
        a = int(a)
        if a < -nd or a >= nd:
            raise ValueError(f'Given axis out of bounds - {a}')
        a = a + nd if a < 0 else a
        if a in out:
            raise ValueError(f'Duplicate axis - {a}')
        out.append(a)
    return tuple(sorted(out))


cdef tuple _reduced_shape(tuple shape, tuple axes, bint keepdims):
    if len(axes) == 0:
        axes = tuple(range(len(shape)))
    if keepdims:
        return tuple([1 if i in axes else s for i, s in enumerate(shape)])
    return tuple([s for i, s in enumerate(shape) if i not in axes])


def _proper_shape(shape):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        return tuple(shape[0])
    return tuple(shape)


# ============================================================================ autodiff node ops
# An Op holds whatever forward saved; bwd(node, adj) returns one gradient ARRAY (or None)
# per input.  A gradient may alias `adj` (add/sub/residual): see Tensor._push_partial.
class Op:
    name = 'op'
    def bwd(self, node, adj):
        raise NotImplementedError


cdef object _unbroadcast(object g, tuple shape):
    """autodiff.pyx:43-101 _make_gradient_compatible on arrays."""
    cdef tuple gs = g.shape
    if gs == shape:
        return g
    cdef int diff = len(gs) - len(shape), i
    axes = tuple([i for i in range(len(gs)) if i < diff or gs[i] != shape[i - diff]])
    data = B.sum(g, axes, None, None, True)
    if len(gs) != len(shape):
        data = B.reshape(data, shape)
    return data


class _EwiseAdd(Op):
    name = 'add'
    def bwd(self, node, adj):  # backward.pyx:60-86 (copies there; aliases here)
        return (adj, adj)


class _ScalarAdd(Op):
    name = 'scalar_add'
    def bwd(self, node, adj):
        return (adj,)


class _Neg(Op):
    name = 'neg'
    def bwd(self, node, adj):
        return (B.negative(adj),)


class _EwiseSub(Op):
    name = 'sub'
    def bwd(self, node, adj):  # backward.pyx:134-172
        x, y = node._inputs
        gy = None
        if y.requires_grad:
            gy = B.negative(adj)
            if node._dtype != y._dtype:
                gy = B.array(gy, y._dtype.name)
        return (adj, gy)


class _ScalarSub(Op):
    name = 'scalar_sub'
    def __init__(self, scalar, commute):
        self.scalar = scalar; self.commute = commute
    def bwd(self, node, adj):  # backward.pyx:175-212
        if not self.commute:
            return (adj,)
        return (B.negative(adj),)


class _EwiseMul(Op):
    name = 'mul'
    def bwd(self, node, adj):  # backward.pyx:215-257
        x, y = node._inputs
        gx = B.multiply(adj, y._data, dtype=x._dtype.name) if x.requires_grad else None
        gy = B.multiply(adj, x._data, dtype=y._dtype.name) if y.requires_grad else None
        return (gx, gy)


class _ScalarMul(Op):
    name = 'scalar_mul'
    def __init__(self, scalar):
        self.scalar = scalar
    def bwd(self, node, adj):  # backward.pyx:259-285
        x = node._inputs[0]
        return (B.multiply(adj, self.scalar, dtype=x._dtype.name),)


class _EwiseDiv(Op):
    name = 'div'
    def bwd(self, node, adj):  # backward.pyx:288-334
        x, y = node._inputs
        gx = B.divide(adj, y._data, dtype=x._dtype.name) if x.requires_grad else None
        gy = None
        if y.requires_grad:
            yd = y._data
            gy = B.negative(B.multiply(adj, B.divide(x._data, B.multiply(yd, yd)), dtype=y._dtype.name))
        return (gx, gy)


class _ScalarDiv(Op):
    name = 'scalar_div'
    def __init__(self, scalar, commute):
        self.scalar = scalar; self.commute = commute
    def bwd(self, node, adj):  # backward.pyx:337-382
        x = node._inputs[0]
        if not self.commute:
            return (B.multiply(adj, 1 / self.scalar, dtype=x._dtype.name),)
        return (B.multiply(adj, B.multiply(-self.scalar, B.power(x._data, -2)), dtype=x._dtype.name),)


class _EwisePow(Op):
    name = 'pow'
    def bwd(self, node, adj):  # backward.pyx:385-439
        x, y = node._inputs
        gx = gy = None
        if x.requires_grad:
            yd = y._data
            gx = B.multiply(adj, B.multiply(yd, B.power(x._data, B.subtract(yd, 1))), dtype=x._dtype.name)
        if y.requires_grad:
            gy = B.multiply(adj, B.multiply(node._data, B.log(x._data)), dtype=y._dtype.name)
        return (gx, gy)


class _ScalarPow(Op):
    name = 'scalar_pow'
    def __init__(self, scalar, commute):
        self.scalar = scalar; self.commute = commute
    def bwd(self, node, adj):  # backward.pyx:442-492
        import math
        x = node._inputs[0]
        if not self.commute:
            s = self.scalar
            return (B.multiply(adj, B.multiply(s, B.power(x._data, s - 1)), dtype=x._dtype.name),)
        return (B.multiply(adj, B.multiply(node._data, math.log(self.scalar)), dtype=x._dtype.name),)


class _BroadcastTo(Op):
    name = 'broadcast_to'
    def bwd(self, node, adj):  # backward.pyx:495-545
        x = node._inputs[0]
        return (_unbroadcast(adj, x.shape),)


class _Sum(Op):
    name = 'sum'
    def __init__(self, axes, keepdims):
        self.axes = axes; self.keepdims = keepdims
    def grad_array(self, node, adj):  # backward.pyx:548-604
        x = node._inputs[0]
        xs = x.shape
        if self.keepdims is True:
            data = B.broadcast_to(adj, xs)
        else:
            ns = node.shape
            bshape = []
            k = 0
            for i in range(len(xs)):
                if len(ns) > 0 and k < len(ns) and xs[i] == ns[k]:
                    bshape.append(xs[i]); k += 1
                else:
                    bshape.append(1)
            data = B.broadcast_to(B.reshape(adj, tuple(bshape)), xs)
        if x._dtype != node._dtype:
            data = B.array(data, x._dtype.name)
        return data
    def bwd(self, node, adj):
        return (self.grad_array(node, adj),)


class _Mean(_Sum):
    name = 'mean'
    def __init__(self, axes, keepdims, observations):
        _Sum.__init__(self, axes, keepdims)
        self.observations = observations
    def bwd(self, node, adj):  # backward.pyx:607-628
        g = self.grad_array(node, adj)
        return (B.multiply(g, 1 / self.observations),)


class _MaxMin(Op):
    name = 'max'
    def __init__(self, axes, keepdims):
        self.axes = axes; self.keepdims = keepdims
    def bwd(self, node, adj):  # backward.pyx:630-701 (quirk Q11: int64 count -> float64 data)
        x = node._inputs[0]
        xd = x._data; nd = node._data; ad = adj
        if self.keepdims is False:
            xs = x.shape; ns = node.shape
            retain = []
            k = 0
            for i in range(len(xs)):
                if len(ns) > 0 and k < len(ns) and xs[i] == ns[k]:
                    retain.append(xs[i]); k += 1
                else:
                    retain.append(1)
            nd = B.reshape(nd, tuple(retain)); ad = B.reshape(ad, tuple(retain))
        mask = B.equal(xd, nd)
        count = B.sum(mask, self.axes if len(self.axes) else None, None, None, True)
        return (B.multiply(B.divide(ad, count), mask),)


class _Matmul(Op):
    name = 'matmul'
    def bwd(self, node, adj):  # backward.pyx:704-742: adj @ y.T, x.T @ adj on .T VIEWS
        x, y = node._inputs
        gx = B.matmul(adj, y._data.T) if x.requires_grad else None
        gy = B.matmul(x._data.T, adj) if y.requires_grad else None
        return (gx, gy)


class _Reshape(Op):
    name = 'reshape'
    def bwd(self, node, adj):  # backward.pyx:744-767
        return (B.reshape(adj, node._inputs[0].shape),)


class _Permute(Op):
    name = 'permute'
    def __init__(self, axes):
        self.axes = axes
    def bwd(self, node, adj):  # backward.pyx:769-803
        n = len(self.axes)
        inv = [0] * n
        for i, a in enumerate(self.axes):
            a = a + n if a < 0 else a
            inv[a] = i
        return (B.transpose(adj, tuple(inv)),)


class _Transpose(Op):
    name = 'transpose'
    def bwd(self, node, adj):  # backward.pyx:806-826
        return (adj.T,)


class _Select(Op):
    name = 'select'
    def __init__(self, idx):
        self.idx = idx
    def bwd(self, node, adj):  # backward.pyx:829-846
        x = node._inputs[0]
        g = B.zeros(x.shape, x._dtype.name)
        g[self.idx] = adj
        return (g,)


class _Relu(Op):
    name = 'relu'
    def bwd(self, node, adj):  # backward.pyx:849-874, one pass instead of greater + multiply
        x = node._inputs[0]
        if x._dtype.name == 'float32' and adj.dtype == np.float32:
            return (B.relu_backward(x._data, adj),)
        return (B.multiply(B.greater(x._data, 0), adj, dtype=node._dtype.name),)


class _Log(Op):
    name = 'log'
    def bwd(self, node, adj):  # backward.pyx:877-900
        return (B.divide(adj, node._inputs[0]._data),)


class _Exp(Op):
    name = 'exp'
    def bwd(self, node, adj):  # backward.pyx:903-926
        return (B.multiply(adj, node._data),)


class _LogSumExp(_Sum):
    name = 'logsumexp'
    def bwd(self, node, adj):  # backward.pyx:929-956
        x = node._inputs[0]
        g = self.grad_array(node, adj)
        axes = self.axes if len(self.axes) else None
        m = B.max(x._data, axes, None, True)
        e = B.exp(B.subtract(x._data, m))
        s = B.sum(e, axes, None, None, True)
        return (B.divide(B.multiply(g, e), s),)


# ============================================================================ Tensor
cdef class Tensor:
    """soket.Tensor (soket/tensor/tensor.pyx:353-2461) on a resident device array.
    Fields (engine.pxd): _data ndarray, _dtype DType, _device, _grad Tensor|None,
    _requires_grad, _retain_grad, _op Op|None (leaf), _inputs, and the backward
    bookkeeping _partials [(array, owned)], _pending, _visit."""

    def __init__(self, array, device=None, dtype=None, requires_grad=None):
        cdef ndarray data
        cdef Tensor other
        self._grad = None
        self._op = None
        self._inputs = ()
        self._partials = None
        self._retain_grad = False
        self._grad_buf = None
        self._lz = None
        self._lshape = None
        self._nuse = 0
        self._device = _default_device() if device is None else device
        if isinstance(array, Tensor):
            other = <Tensor> array
            dt = other._dtype if dtype is None else _dt(dtype)
            data = B.array(other._data, dt.name)
        elif isinstance(array, ndarray):
            data = B.array(array, None if dtype is None else _dt(dtype).name)
            dt = _DTYPES[str(data.dtype)]
        elif isinstance(array, np.ndarray):
            data = B.array(array, None if dtype is None else _dt(dtype).name)
            dt = _DTYPES[str(data.dtype)]
        elif type(array) in (list, tuple, int, float, bool):
            if dtype is None:
                dt = _scalar_dtype(array) if type(array) in (int, float, bool) else _default_dtype
            else:
                dt = _dt(dtype)
            data = B.array(array, dt.name)
        else:
            raise ValueError(f'Unsupported input type: {type(array)}')
        self._data = data
        self._dtype = dt
        self._requires_grad = bool(requires_grad) if requires_grad is not None else False

    # ---- construction helpers ------------------------------------------------------
    @staticmethod
    def _const(ndarray data, dtype=None, requires_grad=False):
        """tensor.pyx:988-1015 _make_const: wrap an array without copying."""
        cdef Tensor t = Tensor.__new__(Tensor)
        t._data = data
        t._dtype = _DTYPES[str(data.dtype)] if dtype is None else dtype
        t._device = _default_device()
        t._grad = None
        t._op = None
        t._inputs = ()
        t._partials = None
        t._requires_grad = requires_grad
        t._retain_grad = False
        t._grad_buf = None
        t._lz = None
        t._lshape = None
        t._nuse = 0
        return t

    # ---- deferred (lazy-mode) nodes --------------------------------------------------
    @property
    def _data(self):
        """The device array; a deferred node is evaluated here, with its whole pending chain."""
        if self._d is None and self._lz is not None:
            _realize(self)
        return self._d

    @_data.setter
    def _data(self, value):
        self._d = value
        self._lz = None

    @staticmethod
    def _deferred(op, tuple inputs, tuple spec, tuple shape, dtype):
        """tensor.pyx:1018-1070 _make_from_op in lazy mode: the node exists (shape, dtype, graph
        links), its data does not yet."""
        cdef Tensor t = Tensor.__new__(Tensor)
        cdef Tensor i
        t._d = None
        t._dtype = dtype
        t._device = _default_device()
        t._grad = None
        t._partials = None
        t._retain_grad = False
        t._grad_buf = None
        t._nuse = 0
        t._lz = spec
        t._lshape = shape
        t._requires_grad = False
        t._op = None
        t._inputs = ()
        for x in inputs:
            if x is not None:
                i = <Tensor> x
                if i._requires_grad:
                    t._requires_grad = True
                if i._d is None and i._lz is not None:
                    i._nuse += 1
        if t._requires_grad:
            t._op = op
            t._inputs = inputs
        return t

    @staticmethod
    def _from_op(op, tuple inputs, ndarray data, dtype):
        """tensor.pyx:1018-1070 _make_from_op (eager): graph links are kept only when
        some input requires grad."""
        cdef Tensor t = Tensor._const(data, dtype, False)
        cdef Tensor i
        for x in inputs:
            if x is not None and (<Tensor> x)._requires_grad:
                t._requires_grad = True
                break
        if t._requires_grad:
            t._op = op
            t._inputs = inputs
        return t

    @staticmethod
    def from_numpy(array):
        return Tensor(array)

    # ---- properties ------------------------------------------------------------------
    @property
    def dtype(self): return self._dtype
    @property
    def shape(self): return self._d.shape if self._d is not None else self._lshape
    @property
    def size(self):
        if self._d is not None:
            return self._d.size
        n = 1
        for s in self._lshape:
            n *= s
        return n
    @property
    def ndim(self): return self._d.ndim if self._d is not None else len(self._lshape)
    @property
    def device(self): return self._device
    @property
    def grad(self): return self._grad
    @property
    def T(self): return self.transpose()
    @property
    def data(self): return self.detach()
    @data.setter
    def data(self, Tensor value):
        # tensor.pyx:967-981 _set_data: rebinding, shape/dtype follow the new data
        self._data = value._data
        self._dtype = value._dtype
        self._device = value._device
    @property
    def requires_grad(self): return self._requires_grad
    @requires_grad.setter
    def requires_grad(self, bint mode):
        if self._op is not None:
            raise RuntimeError('Only detached/leaf tensors can requires_grad!')
        self._requires_grad = mode

    # ---- methods -----------------------------------------------------------------------
    def to(self, device):
        return self

    def copy(self, requires_grad=False):
        return Tensor._const(B.copy(self._data), self._dtype, bool(requires_grad))

    def detach(self):
        """tensor.pyx:950-960: aliases storage."""
        return Tensor._const(self._data, self._dtype, False)

    def numpy(self):
        """Device -> host copy (extension: the reference has no accessor)."""
        return B.asnumpy(self._data)

    def item(self):
        if self._data.ndim != 0:
            raise ValueError('Can only call Tensor.item() on scalar tensors')
        return self._data.item()

    def retain_grad(self):
        self._retain_grad = True

    def backward(self, adj=None):
        if not self._requires_grad:
            raise TypeError('Can call backward() only on tensors with requires_grad=True')
        if adj is not None:
            if not isinstance(adj, Tensor):
                raise TypeError('adjoint must be a Tensor')
            if self._dtype != (<Tensor> adj)._dtype:
                raise RuntimeError('Incompatible adjoint datatype!')
            if self.shape != adj.shape:
                raise ValueError('Incompatible adjoint shape!')
            seed = (<Tensor> adj)._data
        else:
            seed = B.ones(self.shape, self._dtype.name)
        _compute_gradient(self, seed)

    # reductions / views
    def broadcast_to(self, *shape):
        shape = _proper_shape(shape)
        cdef tuple s = self.shape
        if len(shape) < len(s):     # tensor.pyx:1638-1642
            raise RuntimeError('Attempt to reduce a tensor using broadcast_to()')
        for d in shape:
            if type(d) is not int:
                raise TypeError('Expected broadcast shape to be integers!')
        for i in range(1, len(s) + 1):
            if s[-i] != shape[-i] and s[-i] != 1:
                raise RuntimeError('Incompatible broadcast shapes!')   # tensor.pyx:1664-1666
        return Tensor._from_op(_BroadcastTo(), (self,), B.broadcast_to(self._data, shape), self._dtype)

    def sum(self, *axes, dtype=None, keepdims=False):
        ax = _norm_axes(_proper_shape(axes), self._data.ndim)
        dt = self._dtype if dtype is None else _dt(dtype)
        full = len(ax) == 0 or len(_reduced_shape(self.shape, ax, False)) == 0
        data = B.sum(self._data, None if full and not keepdims else (ax if len(ax) else None), dt.name, None, keepdims)
        return Tensor._from_op(_Sum(ax, keepdims), (self,), data, dt)

    def mean(self, *axes, dtype=None, keepdims=False):
        ax = _norm_axes(_proper_shape(axes), self._data.ndim)
        dt = self._dtype if dtype is None else _dt(dtype)
        data = B.mean(self._data, ax if len(ax) else None, dt.name, None, keepdims)
        # quirk Q5 (tensor.pyx:1725-1735): observations compare shapes POSITION-WISE over
        # the result rank, so full reductions and keepdims=False reductions of
        # non-leading axes are not scaled in backward
        cdef tuple s = self.shape
        cdef tuple o = data.shape
        cdef long obs = 1
        for i in range(len(o)):
            if s[i] != o[i]:
                obs *= s[i]
        return Tensor._from_op(_Mean(ax, keepdims, obs), (self,), data, dt)

    def max(self, *axes, dtype=None, keepdims=False):
        ax = _norm_axes(_proper_shape(axes), self._data.ndim)
        data = B.max(self._data, ax if len(ax) else None, None, keepdims)
        return Tensor._from_op(_MaxMin(ax, keepdims), (self,), data, self._dtype)

    def min(self, *axes, dtype=None, keepdims=False):
        ax = _norm_axes(_proper_shape(axes), self._data.ndim)
        data = B.min(self._data, ax if len(ax) else None, None, keepdims)
        return Tensor._from_op(_MaxMin(ax, keepdims), (self,), data, self._dtype)

    def reshape(self, *shape):
        shape = _proper_shape(shape)
        cdef long n = 1
        for s in shape:
            n *= s
        if n != self._data.size:
            raise ValueError('Incompatible shape for reshape!')
        return Tensor._from_op(_Reshape(), (self,), B.reshape(self._data, shape), self._dtype)

    def permute(self, *axes):
        axes = _proper_shape(axes)
        if len(axes) != self._data.ndim:
            raise ValueError('Invalid permutation axes!')
        return Tensor._from_op(_Permute(tuple(axes)), (self,), B.transpose(self._data, axes), self._dtype)

    def transpose(self):
        """quirk Q7 (tensor.pyx:1964-1967): reverses ALL axes."""
        return Tensor._from_op(_Transpose(), (self,), self._data.T, self._dtype)

    def argmax(self, axis=None, keepdims=False):
        if axis is not None and (axis < -self._data.ndim or axis >= self._data.ndim):
            raise ValueError(f'Give axis out of bounds - {axis}')
        return Tensor._const(B.array(B.argmax(self._data, axis, keepdims=keepdims), 'int32'), int32)

    def argmin(self, axis=None, keepdims=False):
        if axis is not None and (axis < -self._data.ndim or axis >= self._data.ndim):
            raise ValueError(f'Give axis out of bounds - {axis}')
        return Tensor._const(B.array(B.argmin(self._data, axis, keepdims=keepdims), 'int32'), int32)

    # ---- arithmetic --------------------------------------------------------------------
    def _binary(self, other, fn, ew_op, sc_op, bint commute=False):
        cdef Tensor t
        if isinstance(other, Tensor):
            t = <Tensor> other
            dt = self._dtype if self._dtype == t._dtype else promote_types(self._dtype, t._dtype)
            bs = _bshape(self.shape, t.shape)
            if _LAZY_STATE and fn in _FUSED_BIN:
                a, b = (t, self) if commute else (self, t)
                if _deferrable(a, bs) and _deferrable(b, bs):
                    return Tensor._deferred(ew_op(), (a, b), ('bin', _FUSED_BIN[fn], (a, b), None, 0), bs, dt)
            if commute:
                return Tensor._from_op(ew_op(), (t, self), fn(t._data, self._data, dtype=dt.name), dt)
            return Tensor._from_op(ew_op(), (self, t), fn(self._data, t._data, dtype=dt.name), dt)
        if type(other) not in (int, float, bool):
            raise RuntimeError(f'Usupported scalar - {other}')
        dt = _scalar_dtype(other)
        if self._dtype != dt:
            dt = promote_types(self._dtype, dt)
        if _LAZY_STATE and fn in _FUSED_BIN and dt.name == 'float32' and _deferrable(self, self.shape):
            return Tensor._deferred(sc_op(other, bool(commute)), (self,),
                                    ('sc', _FUSED_BIN[fn], (self,), float(other), 1 if commute else 0), self.shape, dt)
        if commute:
            return Tensor._from_op(sc_op(other, True), (self,), fn(other, self._data, dtype=dt.name), dt)
        return Tensor._from_op(sc_op(other, False), (self,), fn(self._data, other, dtype=dt.name), dt)

    def __add__(self, other):
        return self._binary(other, B.add, _EwiseAdd, lambda s, c: _ScalarAdd())
    def __radd__(self, other):
        return self._binary(other, B.add, _EwiseAdd, lambda s, c: _ScalarAdd())
    def __sub__(self, other):
        return self._binary(other, B.subtract, _EwiseSub, _ScalarSub)
    def __rsub__(self, other):
        return self._binary(other, B.subtract, _EwiseSub, _ScalarSub, True)
    def __mul__(self, other):
        return self._binary(other, B.multiply, _EwiseMul, lambda s, c: _ScalarMul(s))
    def __rmul__(self, other):
        return self._binary(other, B.multiply, _EwiseMul, lambda s, c: _ScalarMul(s))
    def __truediv__(self, other):
        return self._binary(other, B.divide, _EwiseDiv, _ScalarDiv)
    def __rtruediv__(self, other):
        return self._binary(other, B.divide, _EwiseDiv, _ScalarDiv, True)
    def __pow__(self, other):
        return self._binary(other, B.power, _EwisePow, _ScalarPow)
    def __rpow__(self, other):
        return self._binary(other, B.power, _EwisePow, _ScalarPow, True)
    def __neg__(self):
        if _LAZY_STATE and _deferrable(self, self.shape):
            return Tensor._deferred(_Neg(), (self,), ('un', SK_UOP_NEG, (self,), None, 0), self.shape, self._dtype)
        return Tensor._from_op(_Neg(), (self,), B.negative(self._data), self._dtype)

    def __matmul__(self, other):
        if not isinstance(other, Tensor):
            raise RuntimeError('matmul expects a tensor')
        cdef Tensor t = <Tensor> other
        if self._data.ndim < 2 or t._data.ndim < 2:
            raise RuntimeError('Both tensors must be atleast 2D for matmul!')
        if self.shape[-1] != t.shape[-2]:
            raise RuntimeError('Incompatible shapes for matmul!')
        dt = self._dtype if self._dtype == t._dtype else promote_types(self._dtype, t._dtype)
        _bshape(self.shape, t.shape, 2)
        return Tensor._from_op(_Matmul(), (self, t), B.matmul(self._data, t._data, dtype=dt.name), dt)

    # ---- comparisons: never differentiable; `<=` is `>=` (quirk Q6, tensor.pyx:2448-2451)
    def _cmp(self, other, fn):
        if isinstance(other, Tensor):
            _bshape(self.shape, other.shape)
            return Tensor._const(fn(self._data, (<Tensor> other)._data), bool_)
        if type(other) not in (int, float, bool):
            raise RuntimeError(f'Usupported scalar - {other}')
        return Tensor._const(fn(self._data, other), bool_)
    def __eq__(self, other): return self._cmp(other, B.equal)
    def __ne__(self, other): return self._cmp(other, B.not_equal)
    def __gt__(self, other): return self._cmp(other, B.greater)
    def __ge__(self, other): return self._cmp(other, B.greater_equal)
    def __lt__(self, other): return self._cmp(other, B.less)
    def __le__(self, other): return self._cmp(other, B.greater_equal)
    def __hash__(self): return id(self)

    # ---- indexing ---------------------------------------------------------------------------
    def __getitem__(self, idx):
        if isinstance(idx, Tensor):
            idx = (<Tensor> idx)._data
        # tensor.pyx:1994-2016: integer indices are bounds-checked per leading axis -> ValueError
        cdef tuple s = self.shape
        items = idx if type(idx) is tuple else (idx,)
        for k in range(len(items)):
            if type(items[k]) is int and k < len(s) and (items[k] < -s[k] or items[k] >= s[k]):
                raise ValueError(f'Select index out of bounds - {items[k]}')
        return Tensor._from_op(_Select(idx), (self,), self._data[idx], self._dtype)

    def __setitem__(self, idx, value):
        if self._requires_grad:
            raise RuntimeError('Cannot set value to a parameter!')
        if isinstance(value, Tensor):
            value = (<Tensor> value)._data
        elif type(value) not in (int, float, bool):
            raise RuntimeError(f'Unsupported item - {value}')
        self._data[idx] = value

    def __len__(self):
        return len(self._data)

    def __repr__(self):
        lines = str(self._data).splitlines()
        prefix = ' ' * 13
        res = 'soket.Tensor(' + '\n'.join(lines[:1] + [prefix + l for l in lines[1:]]) + ', dtype=' + self._dtype.name
        res += ', device=GPU:' + str(self._device.id)
        if self._requires_grad:
            res += ', requires_grad=True'
        return res + ')'


# ============================================================================ autodiff engine
_leaf_hook = None


def set_leaf_grad_hook(fn):
    """fn(tensor) is called the moment a LEAF tensor's gradient is final during
    backward (data-parallel all-reduce is issued from here)."""
    global _leaf_hook
    _leaf_hook = fn


cdef int _visit_epoch = 0


cdef inline void _touch(Tensor t, int ep):
    if t._visit != ep:
        t._visit = ep
        t._state = 0
        t._pending = 0
        t._nedges = 0
        t._partials = []


cdef list _topo(Tensor root):
    """autodiff.pyx:170-265: iterative post-order DFS over the inputs that require
    grad (a node may sit on the stack several times; it is expanded once, at its
    first pop, so every consumer finishes after it).  Also counts, per node, how
    many partial adjoints it will receive (one per consuming edge)."""
    global _visit_epoch
    _visit_epoch += 1
    cdef int ep = _visit_epoch
    cdef list order = []
    cdef list stack = [(root, False)]
    cdef Tensor n, i
    _touch(root, ep)
    root._pending = 1
    while stack:
        n, done = stack.pop()
        if done:
            n._state = 2
            order.append(n)
            continue
        if n._state != 0:      # already expanded through another consumer
            continue
        n._state = 1
        stack.append((n, True))
        for x in n._inputs:
            if x is None:
                continue
            i = <Tensor> x
            if not i._requires_grad:
                continue
            _touch(i, ep)
            i._pending += 1
            i._nedges += 1
            if i._state == 0:
                stack.append((i, False))
    return order


cdef inline bint _sole_view(object a):
    """True when `a` is the only array object over its allocation: every view holds a reference to
    the shared Buffer, so a count of one means no other gradient, retained adjoint or forward value
    can observe a write through `a`."""
    cdef Buffer b = (<ndarray> a)._buf      # this local holds one reference itself
    return (<PyObject *> b).ob_refcnt == 2


cdef tuple _sum_partials(list parts):
    """autodiff.pyx:30-41 _sum_nodes, but in place.  Returns (sum, owned).  A partial is written to
    only when it was handed to this node alone (the flag set where it was produced) AND no other
    array shares its storage (`_sole_view`: reshape / transpose / broadcast backward return views
    of the adjoint they were given, add / sub backward hand the same adjoint to both inputs)."""
    cdef object acc, p
    cdef bint owned, o
    acc, owned = parts[0]
    owned = owned and _sole_view(acc)
    for k in range(1, len(parts)):
        p, o = parts[k]
        if (acc.shape == p.shape and str(acc.dtype) == 'float32' and str(p.dtype) == 'float32'
                and acc.is_contiguous and p.is_contiguous):
            if owned:
                F.accumulate_(acc, p)
                continue
            if o and _sole_view(p):
                F.accumulate_(p, acc)
                acc = p; owned = True
                continue
        acc = B.add(acc, p)
        owned = True
    return acc, owned


cdef void _finalize_leaf(Tensor n):
    g = _unbroadcast(_sum_partials(n._partials)[0], n.shape)
    n._partials = None
    # overwritten, never accumulated (autodiff.pyx:221-222).  The dtype TAG is the tensor's own: every
    # backward fn wraps its result with the input's dtype (backward.pyx:14-24), whatever the data is (Q11)
    n._grad = Tensor._const(g, n._dtype, False)
    if _leaf_hook is not None:
        _leaf_hook(n)


cdef void _compute_gradient(Tensor root, object seed):
    """autodiff.pyx:106-164."""
    cdef list order = _topo(root)
    cdef Tensor n, i
    root._partials.append((seed, True))
    root._pending = 0
    cdef int k
    for k in range(len(order) - 1, -1, -1):
        n = <Tensor> order[k]
        if n._op is None:
            if n._partials is not None:      # leaf not yet finalised (no partial arrived)
                if len(n._partials):
                    _finalize_leaf(n)
                else:
                    n._partials = None
            continue
        g0, g_owned = _sum_partials(n._partials)
        g = _unbroadcast(g0, n.shape)
        if g is not g0:
            g_owned = True                   # a fresh reduction
        g0 = None
        n._partials = None
        grads = n._op.bwd(n, g)
        # who receives what: an array handed to several inputs (add / sub backward, add_relu) or kept
        # as this node's retained gradient must not be accumulated into by any of them
        nrecv = {}
        for x, gi in zip(n._inputs, grads):
            if x is not None and gi is not None and (<Tensor> x)._requires_grad:
                nrecv[id(gi)] = nrecv.get(id(gi), 0) + 1
        for x, gi in zip(n._inputs, grads):
            if x is None:
                continue
            i = <Tensor> x
            if not i._requires_grad:
                continue
            i._pending -= 1
            if gi is not None:
                o = nrecv[id(gi)] == 1 and (gi is not g or (g_owned and not n._retain_grad))
                i._partials.append((gi, o))
            if i._pending == 0 and i._op is None and len(i._partials):
                _finalize_leaf(i)
        n._grad = Tensor._const(g, n._dtype, False) if n._retain_grad else None
        g = None; grads = None


# ============================================================================ lazy-mode fusion window
# A deferred node records ONE float32 elementwise step:
#   ('bin', sk_binary_op, (x, y), None, 0)        x (op) y, both tensors
#   ('sc',  sk_binary_op, (x,), scalar, rev)      x (op) scalar, or scalar (op) x when rev
#   ('un',  sk_unary_op,  (x,), None, 0)
# Each operand either has the node's full shape or broadcasts against it as a last-axis vector / a
# single element (the bias add of prototypes.pyx:113, the gamma / beta of forward.pyx:343-352, the
# (B, 1)-free cases); anything else is evaluated eagerly, which also ends the window for its inputs.
_FUSED_BIN = {B.add: SK_OP_ADD, B.subtract: SK_OP_SUB, B.multiply: SK_OP_MUL, B.divide: SK_OP_DIV,
              B.power: SK_OP_POW}
_lazy_stats = {'programs': 0, 'nodes': 0}


def lazy_stats(reset=False):
    """{'programs': fused launches, 'nodes': deferred nodes evaluated by them} since the last reset."""
    out = dict(_lazy_stats)
    if reset:
        _lazy_stats['programs'] = 0
        _lazy_stats['nodes'] = 0
    return out


cdef int _operand_kind(tuple shape, tuple full) except -2:
    """How a tensor of `shape` reads inside a kernel over `full`: SK_F_FULL / SK_F_VECTOR / SK_F_SINGLE,
    or -1 when it does not fit the fused kernel's indexing."""
    cdef long n = 1
    for s in shape:
        n *= s
    if shape == full:
        return SK_F_FULL
    if n == 1:
        return SK_F_SINGLE
    if len(full) >= 1 and len(shape) <= len(full) and shape[-1] == full[-1] and n == full[-1]:
        return SK_F_VECTOR
    return -1


cdef bint _deferrable(Tensor t, tuple full):
    if t._dtype.name != 'float32' or len(full) == 0:
        return False
    if _operand_kind(t.shape, full) < 0:
        return False
    if t._d is not None:
        return t._d._code == SK_F32 and t._d._is_contiguous()
    return t._lz is not None


class _Overflow(Exception):
    pass


cdef class _Program:
    cdef sk_fused_program p
    cdef list keep          # input arrays stay alive until the launch is queued
    cdef dict index
    cdef list free_temps
    cdef tuple full
    cdef int nodes

    def __cinit__(self):
        self.p.n_ops = 0
        self.p.n_in = 0
        self.keep = []
        self.index = {}
        self.free_temps = [2, 1, 0]
        self.nodes = 0

    cdef int op(self, int code, int sub, int src, int idx, int rev, float cst) except -1:
        cdef int k = self.p.n_ops
        if k >= SK_FUSED_MAX_OPS:
            raise _Overflow()
        self.p.code[k] = code; self.p.sub[k] = sub; self.p.src[k] = src; self.p.idx[k] = idx
        self.p.rev[k] = rev; self.p.cst[k] = cst
        self.p.n_ops = k + 1
        return 0

    cdef int input(self, Tensor t) except -1:
        """Index of a MATERIALISED operand (realising it first if it is a deferred node that cannot
        be inlined)."""
        cdef ndarray a = t._data
        key = id(a)
        if key in self.index:
            return self.index[key]
        cdef int k = self.p.n_in
        if k >= SK_FUSED_MAX_INPUTS:
            raise _Overflow()
        if not a._is_contiguous():
            a = a._compact()
        self.p.inp[k] = <const float *> a._ptr
        self.p.in_kind[k] = _operand_kind(a.shape, self.full)
        self.p.n_in = k + 1
        self.keep.append(a)
        self.index[key] = k
        return k

    cdef bint inline(self, Tensor t):
        """A deferred node is evaluated inside this program when it lives in the same index space and
        nobody else is waiting for it; otherwise it is materialised once and read as an input."""
        return t._d is None and t._lz is not None and t._lshape == self.full and t._nuse <= 1

    cdef int emit(self, Tensor t) except -1:
        """Leave the value of `t` in the accumulator."""
        cdef Tensor x, y
        cdef int tmp
        if not self.inline(t):
            self.op(SK_F_LOAD, 0, SK_F_IN, self.input(t), 0, 0.0)
            return 0
        kind, sub, ins, scalar, rev = t._lz
        self.nodes += 1
        if kind == 'un':
            self.emit(<Tensor> ins[0])
            self.op(SK_F_UN, sub, 0, 0, 0, 0.0)
        elif kind == 'sc':
            self.emit(<Tensor> ins[0])
            self.op(SK_F_BIN, sub, SK_F_CONST, 0, rev, <float> scalar)
        else:
            x = <Tensor> ins[0]; y = <Tensor> ins[1]
            if not self.inline(y):
                self.emit(x)
                self.op(SK_F_BIN, sub, SK_F_IN, self.input(y), 0, 0.0)
            elif not self.inline(x):
                self.emit(y)
                self.op(SK_F_BIN, sub, SK_F_IN, self.input(x), 1, 0.0)       # x (op) acc
            else:
                if not self.free_temps:
                    raise _Overflow()
                self.emit(y)
                tmp = self.free_temps.pop()
                self.op(SK_F_STORE, 0, 0, tmp, 0, 0.0)
                self.emit(x)
                self.op(SK_F_BIN, sub, SK_F_TEMP, tmp, 0, 0.0)
                self.free_temps.append(tmp)
        return 0


cdef object _split_point(Tensor t, int depth):
    """A deferred node `depth` levels below `t` along its first pending branch (or the end of that
    branch): materialising it cuts an over-long chain into programs of about that many steps."""
    cdef Tensor cur = t, x
    cdef int i
    for i in range(depth):
        nxt = None
        for obj in cur._lz[2]:
            x = <Tensor> obj
            if x._d is None and x._lz is not None:
                nxt = x
                break
        if nxt is None:
            break
        cur = <Tensor> nxt
    return None if cur is t else cur


cdef int _realize(Tensor t) except -1:
    """Evaluate a deferred node: compile the pending elementwise subgraph hanging off it into one
    accumulator program (tensor.pyx:790-810 does the same walk, calling the backend once per node)
    and launch it."""
    cdef _Program prog
    cdef ndarray out
    cdef int64_t shp[8]
    cdef int i, nd = len(t._lshape)
    while True:
        prog = _Program()
        prog.full = t._lshape
        try:
            # the root is always computed here, whatever its consumer count
            saved = t._nuse
            t._nuse = 0
            try:
                prog.emit(t)
            finally:
                t._nuse = saved
            break
        except _Overflow:
            victim = _split_point(t, 40)
            if victim is None:
                raise RuntimeError('lazy evaluation: an elementwise node does not fit one fused program')
            _realize(<Tensor> victim)
    for i in range(nd):
        shp[i] = t._lshape[i]
    out = _new_array(nd, shp, SK_F32)
    prog.p.n = out._numel()
    prog.p.cols = t._lshape[nd - 1] if nd else 1
    _check(sk_ewise_fused(&prog.p, <float *> out._ptr))
    _lazy_stats['programs'] += 1
    _lazy_stats['nodes'] += prog.nodes
    t._d = out
    t._lz = None
    return 0


# ============================================================================ creation fns
def _mk(ndarray data, dtype, requires_grad):
    return Tensor._const(data, dtype, bool(requires_grad))


def rand(*shape, low=0.0, high=1.0, device=None, dtype=None, requires_grad=False):
    """soket/tensor/creation.pyx:40-78 via Device._rand (device.pyx:204-209)."""
    dt = _dt(dtype)
    return _mk(B.random.uniform(low, high, _proper_shape(shape)).astype(dt.name), dt, requires_grad)


def randn(*shape, mean=0.0, std=1.0, device=None, dtype=None, requires_grad=False):
    dt = _dt(dtype)
    return _mk(B.random.normal(mean, std, _proper_shape(shape)).astype(dt.name), dt, requires_grad)


def randb(*shape, p=0.5, device=None, dtype=None, requires_grad=False):
    dt = _dt(dtype)
    return _mk(B.random.binomial(1, p, _proper_shape(shape)).astype(dt.name), dt, requires_grad)


def zeros(*shape, device=None, dtype=None, requires_grad=False):
    dt = _dt(dtype)
    return _mk(B.zeros(_proper_shape(shape), dt.name), dt, requires_grad)


def ones(*shape, device=None, dtype=None, requires_grad=False):
    dt = _dt(dtype)
    return _mk(B.ones(_proper_shape(shape), dt.name), dt, requires_grad)


def empty(*shape, device=None, dtype=None, requires_grad=False):
    dt = _dt(dtype)
    return _mk(B.empty(_proper_shape(shape), dt.name), dt, requires_grad)


def full(*shape, fill=0.0, device=None, dtype=None, requires_grad=False):
    dt = _dt(dtype)
    return _mk(B.full(_proper_shape(shape), fill, dt.name), dt, requires_grad)


def one_hot(Tensor t, num_classes=-1, device=None, dtype=None, requires_grad=False):
    """creation.pyx:365-418: eye(C)[labels]; num_classes=-1 -> max+1."""
    dt = _dt(dtype)
    if num_classes == -1:
        num_classes = B.max(t._data).item() + 1
    return _mk(B.eye(num_classes, None, 0, dt.name)[t._data], dt, requires_grad)


def zeros_like(Tensor t, device=None, dtype=None, requires_grad=False):
    dt = t._dtype if dtype is None else _dt(dtype)
    return _mk(B.zeros(t.shape, dt.name), dt, requires_grad)


def ones_like(Tensor t, device=None, dtype=None, requires_grad=False):
    dt = t._dtype if dtype is None else _dt(dtype)
    return _mk(B.ones(t.shape, dt.name), dt, requires_grad)


one_like = ones_like      # the reference exports it under this name (creation.pyx:289)


def empty_like(Tensor t, device=None, dtype=None, requires_grad=False):
    dt = t._dtype if dtype is None else _dt(dtype)
    return _mk(B.empty(t.shape, dt.name), dt, requires_grad)


def rand_like(Tensor t, low=0.0, high=1.0, device=None, dtype=None, requires_grad=False):
    return rand(t.shape, low=low, high=high, dtype=t._dtype if dtype is None else dtype, requires_grad=requires_grad)


def randn_like(Tensor t, mean=0.0, std=1.0, device=None, dtype=None, requires_grad=False):
    return randn(t.shape, mean=mean, std=std, dtype=t._dtype if dtype is None else dtype, requires_grad=requires_grad)


def stack(tensors, axis=0):
    """soket/tensor/util.pyx:8-40: no backward."""
    return Tensor._const(B.stack([(<Tensor> t)._data for t in tensors], axis=axis), None, False)


# soket/tensor/detached.pyx:8-79
def log(Tensor x):
    if _LAZY_STATE and _deferrable(x, x.shape):
        return Tensor._deferred(_Log(), (x,), ('un', SK_UOP_LOG, (x,), None, 0), x.shape, x._dtype)
    return Tensor._from_op(_Log(), (x,), B.log(x._data), x._dtype)


def exp(Tensor x):
    if _LAZY_STATE and _deferrable(x, x.shape):
        return Tensor._deferred(_Exp(), (x,), ('un', SK_UOP_EXP, (x,), None, 0), x.shape, x._dtype)
    return Tensor._from_op(_Exp(), (x,), B.exp(x._data), x._dtype)


def logsumexp(Tensor x, *axes, keepdims=False):
    ax = _norm_axes(_proper_shape(axes), x._data.ndim)
    a = ax if len(ax) else None
    m = B.max(x._data, a, None, True)
    res = B.add(B.log(B.sum(B.exp(B.subtract(x._data, m)), a, None, None, True)), m)
    if keepdims is not True:
        res = B.squeeze(res, a)
    return Tensor._from_op(_LogSumExp(ax, keepdims), (x,), res, x._dtype)


# ============================================================================ fused ops
_PRESPLIT = _os_env.environ.get('SOKET_B200_PRESPLIT', '1') != '0'


def set_presplit(on):
    """Linear layers on PRE-SPLIT GEMM operands (default on; SOKET_B200_PRESPLIT=0): each matrix of a
    step is rewritten as fp16 hi / lo once -- the activations by the LayerNorm kernel that produces
    them, a weight once per optimizer step, an adjoint once for dX and dW -- instead of inside every
    GEMM call.  Off = every sk_linear_* call splits its own operands (round-1 behaviour)."""
    global _PRESPLIT
    _PRESPLIT = bool(on)


def presplit_enabled():
    return _PRESPLIT


cdef inline bint _presplit_shapes(ndarray xd, ndarray wd):
    """All three GEMMs of the layer (forward, dX, dW) fit the CTA-pair tcgen05 kernel."""
    if xd._ndim != 2 or wd._ndim != 2 or xd._code != SK_F32 or wd._code != SK_F32:
        return False
    cdef int64_t Bn = xd._shape[0], I = xd._shape[1], O = wd._shape[1]
    return Bn >= 256 and I >= 256 and O >= 128 and I % 8 == 0 and O % 8 == 0 and wd._is_contiguous()


def weight_split_eligible(w):
    """Would `linear` consume this weight array through the pre-split path (for some batch >= 256)?"""
    cdef ndarray wd = <ndarray> w
    return (wd._ndim == 2 and wd._code == SK_F32 and wd._shape[0] >= 256 and wd._shape[1] >= 128
            and wd._shape[0] % 8 == 0 and wd._shape[1] % 8 == 0 and wd._is_contiguous())


cdef object _take_sole_partial(Tensor x):
    """The one partial adjoint x has received so far, if the caller delivers the LAST one and may add
    into it (the block input of model.py:17-37: the residual branch's adjoint arrives first, then
    Linear1's dX) -- the dX GEMM then accumulates in its epilogue instead of a separate pass."""
    if x._partials is None or len(x._partials) != 1 or x._pending != 1:
        return None
    p, o = x._partials[0]
    if not o or not _sole_view(p):
        return None
    cdef ndarray a = <ndarray> p
    if a._code != SK_F32 or a._ndim != 2 or not a._is_contiguous() or a.shape != x.shape:
        return None
    x._partials.pop()
    return p


cdef inline object _slot(object t):
    """The gradient-arena slot of a LEAF parameter (or None).  Only leaves: a non-leaf's adjoint may
    be consumed by reference, and a slot is rewritten on the next backward."""
    if t is None or (<Tensor> t)._op is not None or (<Tensor> t)._nedges != 1:
        return None      # a parameter with several consumers gets its partials summed into a fresh array
    return (<Tensor> t)._grad_buf


class _LinearOp(Op):
    """relu?(X @ W + b) in one GEMM launch (prototypes.pyx:108-115 + :302).  Backward:
    dZ = relu mask * adj (one pass), dX = dZ @ W.T, dW = X.T @ dZ (both on .T views,
    backward.pyx:720-736), db = column sum (autodiff.pyx:43-101 done eagerly)."""
    name = 'linear'
    def __init__(self, relu, xs=None):
        self.relu = relu
        self.xs = xs          # the input's SplitMat (pre-split path): reused by dW = X.T @ adj
    def bwd(self, node, adj):
        x, w, b = node._inputs
        if self.xs is not None:
            return self._bwd_presplit(node, adj)
        if self.relu:
            adj = B.relu_backward(node._data, adj)   # y > 0  <=>  pre-activation > 0
        elif not adj.is_contiguous:
            adj = B.ascontiguousarray(adj)
        a2 = adj if adj.ndim == 2 else B.reshape(adj, (-1, adj.shape[-1]))
        x2 = x._data if x._data.ndim == 2 else B.reshape(x._data, (-1, x.shape[-1]))
        gx = None
        want_b = b is not None and b.requires_grad
        if x.requires_grad and w.requires_grad:
            # _grad_buf: a data-parallel gradient-arena slot the result is written to directly
            if want_b:     # one split of adj shared by both GEMMs, its column sums = the bias gradient
                gx, gw, gb = B.linear_bwd(a2, x2, w._data, True, _slot(w), _slot(b))
            else:
                gx, gw = B.linear_bwd(a2, x2, w._data, False, _slot(w))
                gb = None
            if x._data.ndim != 2:
                gx = B.reshape(gx, x.shape)
            return (gx, gw, gb)
        else:
            if x.requires_grad:
                gx = B.matmul(a2, w._data.T)
                if x._data.ndim != 2:
                    gx = B.reshape(gx, x.shape)
            gw = B.matmul(x2.T, a2) if w.requires_grad else None
        gb = F.colsum(a2) if (b is not None and b.requires_grad) else None
        return (gx, gw, gb)


    def _bwd_presplit(self, node, adj):
        """backward.pyx:704-742 on shared splits: adj is split ONCE (its column sums = the bias
        gradient, autodiff.pyx:84), dX = adj @ W.T reuses the forward's weight split, dW = X.T @ adj
        the forward's input split."""
        x, w, b = node._inputs
        if self.relu:
            adj = B.relu_backward(node._data, adj)
        elif not adj.is_contiguous:
            adj = B.ascontiguousarray(adj)
        gb = None
        if b is not None and b.requires_grad:
            asp, gb = B.split_f16(adj, True, _slot(b))
        else:
            asp = B.get_split(adj)
        gx = gw = None
        if x.requires_grad:
            acc = _take_sole_partial(<Tensor> x)
            gx = B.gemm_split(asp, False, B.get_split(w._data), True, None, False, acc, acc is not None)
        if w.requires_grad:
            gw = B.gemm_split(self.xs, True, asp, False, None, False, _slot(w), False)
        return (gx, gw, gb)


def linear(Tensor x, Tensor w, b=None, bint relu=False):
    if x._data.ndim < 2 or w._data.ndim != 2:
        raise RuntimeError('Both tensors must be atleast 2D for matmul!')
    if x.shape[-1] != w.shape[0]:
        raise RuntimeError('Incompatible shapes for matmul!')
    cdef ndarray xc
    if _PRESPLIT and _presplit_shapes(x._data, w._data) and (b is None or (<Tensor> b)._dtype.name == 'float32'):
        xc = x._data if x._data.is_contiguous else B.ascontiguousarray(x._data)
        xs = B.get_split(xc)
        y = B.gemm_split(xs, False, B.get_split(w._data), False,
                         None if b is None else (<Tensor> b)._data, relu)
        return Tensor._from_op(_LinearOp(relu, xs), (x, w, b), y, x._dtype)
    xd = x._data if x._data.ndim == 2 else B.reshape(x._data, (-1, x.shape[-1]))
    y = B.linear(xd, w._data, None if b is None else (<Tensor> b)._data, relu)
    if x._data.ndim != 2:
        y = B.reshape(y, x.shape[:-1] + (w.shape[1],))
    return Tensor._from_op(_LinearOp(relu), (x, w, b), y, x._dtype)


_FUSE_DROPOUT = _os_env.environ.get('SOKET_B200_FUSE_DROPOUT', '1') != '0'


def set_dropout_fusion(on):
    global _FUSE_DROPOUT
    _FUSE_DROPOUT = bool(on)


class _LayerNormOp(Op):
    """Fused LayerNorm (+ residual add) (+ ReLU): functional.pyx:82-144 +
    forward.pyx:274-353 in one pass; backward.pyx:1025-1132 in one pass."""
    name = 'layer_norm'
    def __init__(self, mean, rstd, relu, has_residual, drop=None):
        self.mean = mean; self.rstd = rstd; self.relu = relu; self.has_residual = has_residual
        self.drop = drop      # (keep, seed): the output went through a fused Dropout
    def bwd(self, node, adj):
        g, b, x, res = node._inputs
        if not adj.is_contiguous:
            adj = B.ascontiguousarray(adj)
        if self.drop is not None:
            keep, seed = self.drop
            want_p = (g is not None and g.requires_grad) or (b is not None and b.requires_grad)
            dx, dg, db = F.layernorm_dropout_bwd(
                adj, x._data, None if g is None else g._data, None if b is None else b._data,
                self.mean, self.rstd, self.relu, keep, 1.0 / keep, seed, want_p, _slot(g), _slot(b), _PRESPLIT)
            return (dg, db, dx if x.requires_grad else None, None)
        mode = 0
        if self.relu:
            mode = 2 if self.has_residual else 1
        want_res = self.has_residual and res.requires_grad
        want_p = (g is not None and g.requires_grad) or (b is not None and b.requires_grad)
        dx, dg, db, dres = F.layernorm_bwd(
            adj, x._data, None if g is None else g._data, None if b is None else b._data,
            self.mean, self.rstd, node._data if mode == 2 else None, mode,
            want_res and mode == 2, want_p, _slot(g), _slot(b), _PRESPLIT)
        if want_res and mode != 2:
            dres = adj    # plain residual add: the adjoint passes through (aliased)
        return (dg, db, dx if x.requires_grad else None, dres if want_res else None)


def layer_norm_dropout(Tensor X, weight, bias, eps, bint relu, keep_rate):
    """dropout([relu](layer_norm(X))) -- the LayerNorm - ReLU - Dropout run of the residual block
    (model.py:24-31) as ONE kernel each way when the fused LayerNorm takes the shape, else the two
    ops in sequence.  SOKET_B200_FUSE_DROPOUT=0 keeps them apart."""
    cdef ndarray xd = X._data
    ok = (_FUSE_DROPOUT and xd.ndim == 2 and X._dtype.name == 'float32' and xd.shape[1] % 4 == 0
          and xd.shape[1] <= 8192 and 0.0 < keep_rate < 1.0)
    for p in (weight, bias):
        if p is not None and ((<Tensor> p)._dtype != X._dtype or (<Tensor> p)._data.size != xd.shape[1]):
            ok = False
    if not ok:
        return dropout(layer_norm(X, weight, bias, eps, relu, None), keep_rate)
    if not xd.is_contiguous:
        xd = B.ascontiguousarray(xd)
    y, mean, rstd, seed = F.layernorm_dropout_fwd(
        xd, None if weight is None else (<Tensor> weight)._data.reshape(-1),
        None if bias is None else (<Tensor> bias)._data.reshape(-1), eps, relu, keep_rate, _PRESPLIT)
    op = _LayerNormOp(mean, rstd, relu, False, (keep_rate, seed))
    return Tensor._from_op(op, (weight, bias, X, None), y, X._dtype)


def layer_norm(Tensor X, weight=None, bias=None, eps=1e-5, bint relu=False, residual=None):
    """soket/nn/functional.pyx:82-144 (+ fused epilogues)."""
    if X._data.ndim < 2:
        raise ValueError('Expected input tensor dimension to be atleast 2D!')
    for p in (weight, bias):
        if p is not None and (<Tensor> p)._dtype != X._dtype:
            raise RuntimeError('Input and parameters should share same datatype!')
    cdef ndarray xd = X._data
    # functional.pyx:108-114: statistics run over ALL non-batch axes and gamma / beta broadcast against
    # the trailing axes; the fused kernel covers the 2-D case (one parameter per normalised element)
    if xd.ndim != 2:
        return _layer_norm_unfused(X, weight, bias, eps, relu, residual)
    if not xd.is_contiguous:
        xd = B.ascontiguousarray(xd)
    cols = xd.shape[1]
    for p in (weight, bias):
        if p is not None and (<Tensor> p)._data.size != cols:
            return _layer_norm_unfused(X, weight, bias, eps, relu, residual)
    if X._dtype.name != 'float32' or cols % 4 != 0 or cols > 8192:
        return _layer_norm_unfused(X, weight, bias, eps, relu, residual)
    rd = None
    rsplit = None
    if residual is not None:
        rd = (<Tensor> residual)._data
        if rd.shape != xd.shape or not rd.is_contiguous:
            rd = B.ascontiguousarray(B.reshape(rd, xd.shape))
        elif _PRESPLIT and isinstance((<ndarray> rd)._meta, SplitMat) and (<SplitMat> (<ndarray> rd)._meta).valid_for(<ndarray> rd):
            rsplit = (<ndarray> rd)._meta     # carries the bound of |residual| the output's scale needs
    y, mean, rstd = F.layernorm_fwd(xd, None if weight is None else (<Tensor> weight)._data.reshape(-1),
                                    None if bias is None else (<Tensor> bias)._data.reshape(-1),
                                    rd, eps, relu, _PRESPLIT and (residual is None or rsplit is not None), rsplit)
    if X._data.ndim != 2:
        y = B.reshape(y, X.shape)
    op = _LayerNormOp(mean, rstd, relu, residual is not None)
    return Tensor._from_op(op, (weight, bias, X, residual), y, X._dtype)


class _NormUnfusedOp(Op):
    """Reference op sequence for shapes the fused kernels do not take."""
    name = 'norm_unfused'
    def __init__(self, axes, obs, xs, r, norm, layernorm):
        self.axes = axes; self.obs = obs; self.xs = xs; self.r = r; self.norm = norm; self.layernorm = layernorm
    def bwd(self, node, adj):
        g, b, x = node._inputs
        xy_axes = (0,) if self.layernorm else self.axes
        dg = db = None
        if g is not None and g.requires_grad:
            dg = B.sum(B.multiply(self.norm, adj), xy_axes, None, None, False)
        if b is not None and b.requires_grad:
            db = B.sum(adj, xy_axes, None, None, False)
        dx = None
        if x.requires_grad:
            ro = 1.0 / self.obs
            if g is None:
                dxn = adj
            elif self.layernorm:
                dxn = B.multiply(adj, g._data)
            else:
                dxn = B.multiply(adj, B.reshape(g._data, self.r.shape))
            xs, r = self.xs, self.r
            dvar = B.sum(B.multiply(B.multiply(dxn, xs), B.multiply(-0.5, B.multiply(B.multiply(r, r), r))),
                         self.axes, None, None, True)
            dmean = B.add(B.sum(B.multiply(dxn, B.multiply(-1.0, r)), self.axes, None, None, True),
                          B.multiply(dvar, B.multiply(ro, B.sum(B.multiply(-2.0, xs), self.axes, None, None, True))))
            dx = B.add(B.multiply(ro, dmean), B.add(B.multiply(dxn, r), B.multiply(B.multiply(dvar, 2.0 * ro), xs)))
        return (dg, db, dx)


def _norm_unfused(Tensor X, weight, bias, tuple axes, long obs, eps, bint layernorm, rm, rv, momentum):
    Z = X._data
    mean = B.mean(Z, axes, None, None, True)
    xs = B.subtract(Z, mean)
    var = B.mean(B.power(xs, 2), axes, None, None, True)
    if rm is not None:
        sub = 1.0 - momentum
        (<Tensor> rm)._data = B.add(B.multiply((<Tensor> rm)._data, sub), B.multiply(mean, momentum))
        (<Tensor> rv)._data = B.add(B.multiply((<Tensor> rv)._data, sub), B.multiply(var, momentum))
    r = B.power(B.add(var, eps), -0.5)
    norm = B.multiply(xs, r)
    out = norm
    if weight is not None:
        gd = (<Tensor> weight)._data
        out = B.multiply(gd if layernorm else B.reshape(gd, mean.shape), norm)
        if bias is not None:
            bd = (<Tensor> bias)._data
            out = B.add(bd if layernorm else B.reshape(bd, mean.shape), out)
    return Tensor._from_op(_NormUnfusedOp(axes, obs, xs, r, norm, layernorm), (weight, bias, X), out, X._dtype)


def _layer_norm_unfused(Tensor X, weight, bias, eps, bint relu, residual):
    axes = tuple(range(1, X._data.ndim))
    obs = 1
    for s in X.shape[1:]:
        obs *= s
    y = _norm_unfused(X, weight, bias, axes, obs, eps, True, None, None, None)
    if residual is not None:
        y = residual + y
    return relu_(y) if relu else y


class _BatchNormOp(Op):
    name = 'batch_norm'
    def __init__(self, mean, rstd, relu):
        self.mean = mean; self.rstd = rstd; self.relu = relu
    def bwd(self, node, adj):
        g, b, x = node._inputs
        if not adj.is_contiguous:
            adj = B.ascontiguousarray(adj)
        dx, dg, db = F.batchnorm_bwd(adj, x._data, None if g is None else g._data,
                                     None if b is None else b._data, self.mean, self.rstd, None,
                                     1 if self.relu else 0)
        return (dg if (g is not None and g.requires_grad) else None,
                db if (b is not None and b.requires_grad) else None,
                dx if x.requires_grad else None)


def batch_norm(Tensor X, running_mean, running_var, gamma=None, beta=None, training=False,
               momentum=0.1, eps=1e-5, bint relu=False):
    """soket/nn/functional.pyx:8-79.  Quirk Q4 (forward.pyx:281): the reference's
    `training` flag is a pointer cast, so batch statistics + running-stat updates are
    ALWAYS used; running stats start 0-d and become (1, C) after the first call."""
    if X._data.ndim < 2:
        raise ValueError('Expected input tensor dimension to be atleast 2D!')
    for p in (running_mean, running_var, gamma, beta):
        if p is not None and (<Tensor> p)._dtype != X._dtype:
            raise RuntimeError('Input and parameters should share same datatype!')
    nd = X._data.ndim
    axes = (0,) + tuple(range(2, nd))
    obs = X.shape[0]
    for s in X.shape[2:]:
        obs *= s
    cols = X.shape[1]
    if nd != 2 or X._dtype.name != 'float32' or cols % 4 != 0 or not X._data.is_contiguous:
        y = _norm_unfused(X, gamma, beta, axes, obs, eps, False, running_mean, running_var, momentum)
        return relu_(y) if relu else y
    rm = rv = None
    if running_mean is not None:
        # materialise the (1, C) running stats the reference ends up with
        for t, fillv in ((running_mean, None), (running_var, None)):
            tt = <Tensor> t
            if tt._data.size != cols:
                tt._data = B.ascontiguousarray(B.broadcast_to(B.reshape(tt._data, (1, 1)), (1, cols)))
        rm = (<Tensor> running_mean)._data
        rv = (<Tensor> running_var)._data
    y, mean, rstd = F.batchnorm_fwd(X._data, None if gamma is None else (<Tensor> gamma)._data,
                                    None if beta is None else (<Tensor> beta)._data,
                                    None if rm is None else rm.reshape(-1),
                                    None if rv is None else rv.reshape(-1), eps, momentum, relu)
    return Tensor._from_op(_BatchNormOp(mean, rstd, relu), (gamma, beta, X), y, X._dtype)


class _AddReluOp(Op):
    name = 'add_relu'
    def bwd(self, node, adj):
        g = B.relu_backward(node._data, adj)
        return (g, g)


def relu_(Tensor x):
    """soket/nn/prototypes.pyx:302-311."""
    if _LAZY_STATE and _deferrable(x, x.shape):
        return Tensor._deferred(_Relu(), (x,), ('un', SK_UOP_RELU, (x,), None, 0), x.shape, x._dtype)
    return Tensor._from_op(_Relu(), (x,), B.maximum(x._data, 0), x._dtype)


def add_relu(Tensor a, Tensor b):
    """relu(a + b): Residual + outer ReLU (prototypes.pyx:272-273, model.py:34-37)."""
    if a.shape != b.shape or a._dtype.name != 'float32' or b._dtype.name != 'float32':
        return relu_(a + b)
    return Tensor._from_op(_AddReluOp(), (a, b),
                           F.add_relu(B.ascontiguousarray(a._data), B.ascontiguousarray(b._data)), a._dtype)


class _DropoutOp(Op):
    name = 'dropout'
    def __init__(self, seed, keep, r_keep):
        self.seed = seed; self.keep = keep; self.r_keep = r_keep
    def bwd(self, node, adj):  # two scalar/elementwise multiplies in the reference; here one pass,
        # the mask regenerated from the forward's seed instead of stored
        if not adj.is_contiguous:
            adj = B.ascontiguousarray(adj)
        return (F.dropout_bwd(adj, self.keep, self.r_keep, self.seed),)


def dropout(Tensor x, keep_rate):
    """prototypes.pyx:746-760: X * Bernoulli(keep) * (1/keep), fused."""
    if x._dtype.name != 'float32':
        m = B.random.binomial(1, keep_rate, x.shape).astype(x._dtype.name)
        return (x * Tensor._const(m)) * (1.0 / keep_rate)
    out, seed = F.dropout_seeded(B.ascontiguousarray(x._data), keep_rate)
    return Tensor._from_op(_DropoutOp(seed, keep_rate, 1.0 / keep_rate), (x,), out, x._dtype)


class _SoftmaxCEOp(Op):
    name = 'softmax_ce'
    def __init__(self, dlogits):
        self.dlogits = dlogits
    def bwd(self, node, adj):
        # forward already produced (softmax - onehot)/B; the seed adjoint of a loss is 1
        return (B.multiply(adj, self.dlogits),)


def softmax_cross_entropy(Tensor X, Tensor targets):
    """Mean softmax-CE of (B, C) logits against integer labels in ONE kernel:
    forward.pyx:250-271 + backward.pyx:959-1022, one-hot never materialised."""
    loss, dl = F.softmax_ce(B.ascontiguousarray(X._data), targets._data, X._requires_grad)
    return Tensor._from_op(_SoftmaxCEOp(dl), (X,), loss, X._dtype)


class _SXEntUnfusedOp(Op):
    """Reference op sequence of the loss (forward.pyx:250-271, backward.pyx:959-1022)
    for reductions / shapes the fused kernel does not take."""
    name = 'sxent_unfused'
    def __init__(self, axes, onehot, reduction):
        self.axes = axes; self.onehot = onehot; self.reduction = reduction
    def bwd(self, node, adj):
        x = node._inputs[0]
        xd = x._data
        m = B.max(xd, self.axes, None, True)
        e = B.exp(B.subtract(xd, m))
        s = B.sum(e, self.axes, None, None, True)
        g = B.subtract(B.divide(e, s), self.onehot, dtype=x._dtype.name)
        if self.reduction == 'mean' and xd.ndim >= 2:
            g = B.multiply(g, 1 / float(x.shape[0]))
        elif self.reduction == 'none' and xd.ndim >= 2:
            adj = B.reshape(adj, (x.shape[0], 1) + tuple(x.shape[2:]))
        return (B.multiply(adj, g),)


def softmax_cross_entropy_unfused(Tensor X, onehot, tuple axes, str reduction):
    xd = X._data
    m = B.max(xd, axes, None, True)
    lse = B.squeeze(B.add(B.log(B.sum(B.exp(B.subtract(xd, m)), axes, None, None, True)), m), axes)
    batch = B.subtract(lse, B.sum(B.multiply(xd, onehot, dtype=X._dtype.name), axes))
    if reduction == 'sum':
        out = B.sum(batch, (0,))
    elif reduction == 'mean':
        out = B.mean(batch, (0,))
    else:
        out = batch
    return Tensor._from_op(_SXEntUnfusedOp(axes, onehot, reduction), (X,), out, X._dtype)
