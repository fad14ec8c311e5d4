# cython: language_level=3, boundscheck=False, wraparound=False, cdivision=True
"""soket_b200._core -- the device array (`ndarray`) and the backend array
interface Soket dispatches to, written in Cython over the C-ABI of
libsoketb200.so (include/soket_b200.h).

This module is what sits in the seam where CuPy sits in the reference
(SURVEY.md section 8b):

  1. the ``Device`` GPU branch's ``_backend`` module surface
     (soket/backend/device.pyx:52-71): ``random.uniform/normal/binomial``,
     ``zeros/ones/eye/empty/full``, ``array``, ``ndarray``, ``asnumpy``;
  2. the 29 callables of the intern table's GPU column
     (soket/tensor/ops/intern.pyx:45-76) with the exact positional / keyword
     conventions the forward/backward ops use (forward.pyx, backward.pyx);
  3. the array-object protocol used on results: ``.shape .size .dtype .T
     .astype .item __getitem__ __setitem__ __str__``.

There is NO CPU fallback: every function needs a CUDA device and raises
``RuntimeError`` (from ``sk_last_error``) otherwise.  NumPy is used for host
metadata only (dtype objects, type promotion, staging buffers for H2D/D2H).
"""
from libc.stdint cimport int64_t, uint32_t, uint64_t, uintptr_t
from libc.string cimport memcpy, memset
from cpython.ref cimport PyObject
cimport cython

from soket_b200._abi cimport *

import numpy as np

cdef object _np = np
cdef object _np_ndarray = np.ndarray
cdef object _np_generic = np.generic


# --------------------------------------------------------------------------- errors
cdef int _check(int rc) except -1:
    if rc != SK_OK:
        msg = sk_last_error().decode('utf-8', 'replace')
        if rc == 5:
            raise MemoryError(msg)
        if rc == 6:
            raise IndexError(msg)
        raise RuntimeError(msg)
    return 0


# --------------------------------------------------------------------------- dtypes
class _BFloat16:
    """Tag for bf16 storage (GEMM operand only; NumPy has no such dtype)."""
    itemsize = 2
    name = 'bfloat16'
    kind = 'V'
    def __str__(self): return 'bfloat16'
    def __repr__(self): return "dtype('bfloat16')"
    def __eq__(self, o): return o is self or str(o) == 'bfloat16'
    def __hash__(self): return hash('bfloat16')

bfloat16 = _BFloat16()

cdef dict _CODE_OF = {}
cdef list _NP_OF = [None] * 13
for _name, _c in (('bool', SK_BOOL), ('int8', SK_I8), ('uint8', SK_U8), ('int16', SK_I16),
                  ('uint16', SK_U16), ('int32', SK_I32), ('uint32', SK_U32), ('int64', SK_I64),
                  ('uint64', SK_U64), ('float16', SK_F16), ('float32', SK_F32),
                  ('float64', SK_F64)):
    _CODE_OF[np.dtype(_name)] = _c
    _NP_OF[_c] = np.dtype(_name)
_NP_OF[SK_BF16] = bfloat16

cdef int _ITEMSIZE[13]
_ITEMSIZE[:] = [1, 1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8, 2]

cdef object _F32 = np.dtype('float32')
cdef object _F64 = np.dtype('float64')
cdef object _I64 = np.dtype('int64')
cdef object _U64 = np.dtype('uint64')
cdef object _BOOL = np.dtype('bool')


cdef int _code(object dt) except -1:
    """numpy dtype / str / python type -> sk_dtype code."""
    if dt is bfloat16 or (isinstance(dt, str) and dt == 'bfloat16'):
        return SK_BF16
    cdef object d = dt if type(dt) is type(_F32) else np.dtype(dt)
    try:
        return <int> _CODE_OF[d]
    except KeyError:
        raise TypeError(f"soket_b200: unsupported dtype '{dt}'")


# --------------------------------------------------------------------------- memory
cdef class Buffer:
    """Owner of one device allocation from the caching allocator."""

    def __cinit__(self):
        self.ptr = 0
        self.nbytes = 0
        self.version = 0
        self.parent = None

    def __dealloc__(self):
        if self.ptr != 0 and self.parent is None:
            sk_free(<void *> self.ptr)
        self.ptr = 0


cdef Buffer _alloc(size_t nbytes):
    cdef Buffer b = Buffer.__new__(Buffer)
    cdef void *p = NULL
    _check(sk_malloc(nbytes, &p))
    b.ptr = <size_t> p
    b.nbytes = nbytes
    return b


def arena_view(ndarray arena, int64_t offset, shape):
    """A C-contiguous array of `shape` over elements [offset, offset + prod(shape)) of the 1-D contiguous
    `arena`, with a write-version of its OWN (data-parallel training keeps every parameter in one
    IPC-exported allocation; an in-place update of one parameter must not invalidate the operand
    splits derived from the others).  The window keeps the arena's allocation alive."""
    cdef ndarray a = ndarray.__new__(ndarray)
    cdef int64_t n = 1
    cdef int i, ndim = len(shape)
    if arena._ndim != 1 or not arena._is_contiguous():
        raise ValueError('arena_view: the arena must be a contiguous 1-D array')
    if ndim > SK_MAX_NDIM:
        raise ValueError(f'soket_b200: at most {SK_MAX_NDIM} dimensions are supported')
    a._ndim = ndim
    for i in range(ndim - 1, -1, -1):
        a._shape[i] = shape[i]
        a._strides[i] = n
        n *= shape[i]
    if offset < 0 or offset + n > arena._shape[0]:
        raise ValueError('arena_view: window outside the arena')
    a._code = arena._code
    a._np_dtype = arena._np_dtype
    cdef Buffer b = Buffer.__new__(Buffer)
    b.ptr = arena._ptr + <size_t> (offset * _ITEMSIZE[arena._code])
    b.nbytes = <size_t> n * _ITEMSIZE[arena._code]
    b.parent = arena._buf
    a._buf = b
    a._ptr = b.ptr
    a._readonly = False
    return a


cdef ndarray _new_array(int ndim, const int64_t *shape, int code):
    """Fresh C-contiguous array."""
    cdef ndarray a = ndarray.__new__(ndarray)
    cdef int64_t n = 1
    cdef int i
    if ndim > SK_MAX_NDIM:
        raise ValueError(f'soket_b200: at most {SK_MAX_NDIM} dimensions are supported')
    a._ndim = ndim
    for i in range(ndim - 1, -1, -1):
        a._shape[i] = shape[i]
        a._strides[i] = n
        n *= shape[i]
    a._code = code
    a._np_dtype = _NP_OF[code]
    a._buf = _alloc(<size_t> n * _ITEMSIZE[code])
    a._ptr = a._buf.ptr
    a._readonly = False
    return a


cdef ndarray _new_like_shape(ndarray ref, int code):
    return _new_array(ref._ndim, ref._shape, code)


cdef ndarray _new_from_tuple(object shape, int code):
    cdef int64_t shp[8]
    cdef int nd, i
    if not isinstance(shape, (tuple, list)):
        shape = (shape,)
    nd = len(shape)
    if nd > SK_MAX_NDIM:
        raise ValueError(f'soket_b200: at most {SK_MAX_NDIM} dimensions are supported')
    for i in range(nd):
        shp[i] = <int64_t> shape[i]
        if shp[i] < 0:
            raise ValueError('negative dimensions are not allowed')
    return _new_array(nd, shp, code)


cdef float *_fptr(ndarray a) except NULL:
    """Raw float32 pointer of a contiguous fp32 array (fused-kernel entry points)."""
    if a._code != SK_F32:
        raise TypeError('expected a float32 device array')
    if not a._is_contiguous():
        raise ValueError('expected a contiguous device array')
    if a._ptr == 0:
        raise ValueError('null device array')
    return <float *> a._ptr


# --------------------------------------------------------------------------- ndarray
@cython.final
cdef class ndarray:
    """Strided view of device memory: pointer + dtype + shape + element strides.

    ``broadcast_to``, ``transpose``/``.T``, ``reshape`` (when possible), basic
    slicing and ``squeeze`` are zero-copy views, as in NumPy -- the reference
    stores such views as gradients (backward.pyx:570,593-596) and feeds ``.T``
    views straight into matmul (backward.pyx:722,734).  Kernels read strided
    operands directly; ``_compact()`` materialises when a kernel needs it.
    """

    def __cinit__(self):
        self._ptr = 0
        self._ndim = 0
        self._code = SK_F32
        self._readonly = False
        self._meta = None

    cdef void _touch(self):
        """The contents were (or are about to be) overwritten in place: everything derived from the
        old contents -- an operand split, a known |max| -- is stale, through every view."""
        self._buf.version += 1
        self._meta = None

    def mark_modified(self):
        self._touch()

    # ---- C-level helpers ------------------------------------------------------
    cdef int64_t _numel(self):
        cdef int64_t n = 1
        cdef int i
        for i in range(self._ndim):
            n *= self._shape[i]
        return n

    cdef bint _is_contiguous(self):
        cdef int64_t expect = 1
        cdef int i
        for i in range(self._ndim - 1, -1, -1):
            if self._shape[i] == 1:
                continue
            if self._strides[i] != expect:
                return False
            expect *= self._shape[i]
        return True

    cdef void _desc(self, sk_array *d):
        cdef int i
        d.data = <void *> self._ptr
        d.dtype = self._code
        d.ndim = self._ndim
        for i in range(self._ndim):
            d.shape[i] = self._shape[i]
            d.strides[i] = self._strides[i]

    cdef int _desc_bcast(self, sk_array *d, int ndim, const int64_t *shape) except -1:
        """Descriptor of `self` broadcast (NumPy rules) to `shape`."""
        cdef int i, off = ndim - self._ndim
        if off < 0:
            raise ValueError('operand has more dimensions than the broadcast shape')
        d.data = <void *> self._ptr
        d.dtype = self._code
        d.ndim = ndim
        for i in range(ndim):
            d.shape[i] = shape[i]
            if i < off:
                d.strides[i] = 0
            elif self._shape[i - off] == shape[i]:
                d.strides[i] = self._strides[i - off] if shape[i] != 1 else 0
            elif self._shape[i - off] == 1:
                d.strides[i] = 0
            else:
                raise ValueError(
                    f'operands could not be broadcast together: {self.shape} -> '
                    f'{tuple([shape[k] for k in range(ndim)])}')
        return 0

    cdef ndarray _view(self, int ndim, const int64_t *shape, const int64_t *strides, int64_t offset):
        cdef ndarray v = ndarray.__new__(ndarray)
        cdef int i
        v._buf = self._buf
        v._ptr = self._ptr + <size_t> (offset * _ITEMSIZE[self._code])
        v._code = self._code
        v._np_dtype = self._np_dtype
        v._ndim = ndim
        v._readonly = self._readonly
        for i in range(ndim):
            v._shape[i] = shape[i]
            v._strides[i] = strides[i]
        return v

    cdef ndarray _compact(self):
        """C-contiguous array with the same contents (self if already contiguous)."""
        if self._is_contiguous():
            return self
        cdef ndarray out = _new_array(self._ndim, self._shape, self._code)
        cdef sk_array s, d
        self._desc(&s)
        out._desc(&d)
        _check(sk_copy(&s, &d))
        return out

    # ---- array protocol used by Soket --------------------------------------------
    @property
    def shape(self):
        cdef int i
        return tuple([self._shape[i] for i in range(self._ndim)])

    @property
    def strides(self):
        """Byte strides, as NumPy reports them."""
        cdef int i
        return tuple([self._strides[i] * _ITEMSIZE[self._code] for i in range(self._ndim)])

    @property
    def ndim(self):
        return self._ndim

    @property
    def size(self):
        return self._numel()

    @property
    def dtype(self):
        return self._np_dtype

    @property
    def itemsize(self):
        return _ITEMSIZE[self._code]

    @property
    def nbytes(self):
        return self._numel() * _ITEMSIZE[self._code]

    @property
    def data_ptr(self):
        return self._ptr

    @property
    def is_contiguous(self):
        return self._is_contiguous()

    @property
    def T(self):
        return _transpose(self, None)

    def __len__(self):
        if self._ndim == 0:
            raise TypeError('len() of unsized object')
        return self._shape[0]

    def get(self):
        """Device -> host copy as a NumPy array."""
        return asnumpy(self)

    def item(self):
        if self._numel() != 1:
            raise ValueError('can only convert an array of size 1 to a Python scalar')
        return asnumpy(self).item()

    def tolist(self):
        return asnumpy(self).tolist()

    def astype(self, dtype, copy=True):
        cdef int code = _code(dtype)
        if code == self._code and not copy:
            return self
        return _cast_copy(self, code)

    def copy(self):
        return _cast_copy(self, self._code)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return reshape(self, shape)

    def transpose(self, *axes):
        if len(axes) == 0:
            return _transpose(self, None)
        if len(axes) == 1 and (axes[0] is None or isinstance(axes[0], (tuple, list))):
            return _transpose(self, axes[0])
        return _transpose(self, axes)

    def squeeze(self, axis=None):
        return squeeze(self, axis)

    def sum(self, axis=None, dtype=None, out=None, keepdims=False):
        return sum(self, axis, dtype, out, keepdims)

    def mean(self, axis=None, dtype=None, out=None, keepdims=False):
        return mean(self, axis, dtype, out, keepdims)

    def max(self, axis=None, out=None, keepdims=False):
        return max(self, axis, out, keepdims)

    def min(self, axis=None, out=None, keepdims=False):
        return min(self, axis, out, keepdims)

    def argmax(self, axis=None, out=None, keepdims=False):
        return argmax(self, axis, out, keepdims=keepdims)

    def argmin(self, axis=None, out=None, keepdims=False):
        return argmin(self, axis, out, keepdims=keepdims)

    def fill(self, value):
        self._touch()
        _fill(self, value)

    def __getitem__(self, idx):
        return _getitem(self, idx)

    def __setitem__(self, idx, value):
        _setitem(self, idx, value)

    def __str__(self):
        return str(asnumpy(self))

    def __repr__(self):
        r = repr(asnumpy(self))
        return 'soket_b200.' + r if r.startswith('array') else r

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __bool__(self):
        if self._numel() != 1:
            raise ValueError('The truth value of an array with more than one element is ambiguous.')
        return bool(self.item())

    # arithmetic dunders (NumPy semantics) -- conveniences; Soket itself goes
    # through the module-level functions
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    def __pow__(self, o): return power(self, o)
    def __rpow__(self, o): return power(o, self)
    def __neg__(self): return negative(self)
    def __matmul__(self, o): return matmul(self, o)
    def __eq__(self, o): return equal(self, o)
    def __ne__(self, o): return not_equal(self, o)
    def __gt__(self, o): return greater(self, o)
    def __ge__(self, o): return greater_equal(self, o)
    def __lt__(self, o): return less(self, o)
    def __le__(self, o): return less_equal(self, o)
    __hash__ = None


# --------------------------------------------------------------------------- H2D / D2H
cdef ndarray _from_numpy(object arr, int code):
    """Upload a NumPy array (cast on the host to `code` first if needed)."""
    cdef object want = _NP_OF[code]
    cdef object host = np.asarray(arr, dtype=want, order='C')
    cdef ndarray out = _new_from_tuple(host.shape, code)
    cdef size_t nbytes = <size_t> host.nbytes
    cdef uintptr_t src
    if nbytes:
        src = <uintptr_t> host.ctypes.data
        _check(sk_h2d(<void *> out._ptr, <const void *> src, nbytes))
    return out


cdef ndarray _as_device(object x):
    if type(x) is ndarray:
        return <ndarray> x
    if isinstance(x, (_np_ndarray, _np_generic)):
        arr = np.asarray(x)
        return _from_numpy(arr, _code(arr.dtype))
    raise TypeError(f'soket_b200: cannot use {type(x).__name__} as a device array operand')


def asnumpy(a, stream=None, order='C'):
    """Device -> host (replaces cupy.asnumpy, soket/tensor/tensor.pyx:226,387)."""
    if isinstance(a, (_np_ndarray, _np_generic)):
        return np.asarray(a)
    cdef ndarray src = _as_device(a)
    cdef ndarray c
    if src._code == SK_BF16:
        c = _cast_copy(src, SK_F32)
    else:
        c = src._compact()
    cdef object host = np.empty(c.shape, dtype=_NP_OF[c._code])
    cdef size_t nbytes = <size_t> host.nbytes
    cdef uintptr_t dst
    if nbytes:
        dst = <uintptr_t> host.ctypes.data
        _check(sk_d2h(<void *> dst, <const void *> c._ptr, nbytes))
    else:
        _check(sk_sync())
    return host


cdef ndarray _cast_copy(ndarray src, int code):
    cdef ndarray out = _new_array(src._ndim, src._shape, code)
    cdef sk_array s, d
    src._desc(&s)
    out._desc(&d)
    if code == SK_BF16 and src._code == SK_F32:
        _check(sk_cast_bf16(&s, &d))
    else:
        _check(sk_copy(&s, &d))
    return out


def array(obj, dtype=None, copy=True, order='K', subok=False, ndmin=0):
    """Replaces cupy.array / np.array at the seam (intern slot _ARRAY,
    Device._backend.array: soket/tensor/tensor.pyx:386-442): converts nested
    lists / Python scalars / NumPy arrays (H2D) and copies / casts device arrays."""
    cdef ndarray src
    cdef int code
    if type(obj) is ndarray:
        src = <ndarray> obj
        code = src._code if dtype is None else _code(dtype)
        return _cast_copy(src, code)
    if isinstance(obj, (_np_ndarray, _np_generic)):
        arr = np.asarray(obj)
        return _from_numpy(arr, _code(arr.dtype) if dtype is None else _code(dtype))
    if isinstance(obj, (list, tuple)) and _contains_device(obj):
        return stack([array(o, dtype) for o in obj])
    host = np.array(obj, dtype=None if dtype is None else _NP_OF[_code(dtype)])
    return _from_numpy(host, _code(host.dtype))


cdef bint _contains_device(object seq):
    for o in seq:
        if type(o) is ndarray:
            return True
    return False


def asarray(obj, dtype=None):
    if type(obj) is ndarray and (dtype is None or _code(dtype) == (<ndarray> obj)._code):
        return obj
    return array(obj, dtype)


def copy(a):
    """np.copy (intern slot _COPY, soket/tensor/tensor.pyx:606)."""
    cdef ndarray src = _as_device(a)
    return _cast_copy(src, src._code)


def to_bf16(a):
    """float32 -> bfloat16 (round-to-nearest-even) operand for matmul(..., SK_MM_BF16)."""
    cdef ndarray src = _as_device(a)
    if src._code != SK_F32:
        raise TypeError('to_bf16 expects a float32 array')
    return _cast_copy(src._compact(), SK_BF16)


# --------------------------------------------------------------------------- creation
cdef int _fill(ndarray a, object value) except -1:
    cdef sk_array d
    a._desc(&d)
    if isinstance(value, (bool, int, np.integer, np.bool_)) :
        _check(sk_fill(&d, 0.0, <int64_t> int(value), 1))
    else:
        _check(sk_fill(&d, <double> float(value), 0, 0))
    return 0


def empty(shape, dtype=float, order='C'):
    return _new_from_tuple(shape, _code(dtype))


def zeros(shape, dtype=float, order='C'):
    cdef ndarray a = _new_from_tuple(shape, _code(dtype))
    _fill(a, 0)
    return a


def ones(shape, dtype=float, order='C'):
    cdef ndarray a = _new_from_tuple(shape, _code(dtype))
    _fill(a, 1)
    return a


def full(shape, fill_value, dtype=None, order='C'):
    if dtype is None:
        dtype = np.result_type(fill_value)
    cdef ndarray a = _new_from_tuple(shape, _code(dtype))
    _fill(a, fill_value)
    return a


def eye(N, M=None, k=0, dtype=float, order='C'):
    """np.eye(N, M, k, dtype) -- Device._one_hot calls eye(C, None, 0, dtype)
    (soket/backend/device.pyx:236-239)."""
    cdef ndarray a = _new_from_tuple((N, N if M is None else M), _code(dtype))
    cdef sk_array d
    a._desc(&d)
    _check(sk_eye(&d, <int64_t> k))
    return a


def zeros_like(a, dtype=None):
    cdef ndarray s = _as_device(a)
    cdef ndarray o = _new_array(s._ndim, s._shape, s._code if dtype is None else _code(dtype))
    _fill(o, 0)
    return o


def ones_like(a, dtype=None):
    cdef ndarray s = _as_device(a)
    cdef ndarray o = _new_array(s._ndim, s._shape, s._code if dtype is None else _code(dtype))
    _fill(o, 1)
    return o


def empty_like(a, dtype=None):
    cdef ndarray s = _as_device(a)
    return _new_array(s._ndim, s._shape, s._code if dtype is None else _code(dtype))


# --------------------------------------------------------------------------- views
cdef ndarray _transpose(ndarray a, object axes):
    cdef int64_t shp[8]
    cdef int64_t st[8]
    cdef int i, ax, nd = a._ndim
    cdef unsigned int seen = 0
    if axes is None:
        for i in range(nd):
            shp[i] = a._shape[nd - 1 - i]
            st[i] = a._strides[nd - 1 - i]
    else:
        if len(axes) != nd:
            raise ValueError("axes don't match array")
        for i in range(nd):
            ax = axes[i]
            if ax < 0:
                ax += nd
            if ax < 0 or ax >= nd:
                raise ValueError(f'axis {axes[i]} is out of bounds for array of dimension {nd}')
            if seen & (1u << ax):
                raise ValueError('repeated axis in transpose')
            seen |= (1u << ax)
            shp[i] = a._shape[ax]
            st[i] = a._strides[ax]
    return a._view(nd, shp, st, 0)


def transpose(a, axes=None):
    """np.transpose (intern slot _TRANSPOSE; forward.pyx:188-199): a VIEW."""
    return _transpose(_as_device(a), axes)


def broadcast_to(a, shape, subok=False):
    """np.broadcast_to (intern slot _BCASTTO; forward.pyx:120-125,
    backward.pyx:570,593): zero-stride read-only VIEW."""
    cdef ndarray s = _as_device(a)
    cdef int64_t shp[8]
    cdef int nd, i
    cdef sk_array d
    if not isinstance(shape, (tuple, list)):
        shape = (shape,)
    nd = len(shape)
    if nd > SK_MAX_NDIM:
        raise ValueError('too many dimensions')
    for i in range(nd):
        shp[i] = shape[i]
    if nd < s._ndim:
        raise ValueError('input operand has more dimensions than allowed by the axis remapping')
    s._desc_bcast(&d, nd, shp)
    cdef ndarray v = s._view(nd, shp, d.strides, 0)
    v._readonly = True
    return v


cdef bint _nocopy_reshape(ndarray a, int new_nd, const int64_t *new_shape, int64_t *new_strides):
    """NumPy's no-copy reshape rule: returns True and fills new_strides when the
    reshape can be expressed as a view."""
    cdef int64_t oshape[8]
    cdef int64_t ostride[8]
    cdef int old_nd = 0, i, oi, oj, ni, nj, ok
    cdef int64_t np_, op
    for i in range(a._ndim):
        if a._shape[i] != 1:
            oshape[old_nd] = a._shape[i]
            ostride[old_nd] = a._strides[i]
            old_nd += 1
    oi = 0; oj = 1; ni = 0; nj = 1
    while ni < new_nd and oi < old_nd:
        np_ = new_shape[ni]
        op = oshape[oi]
        while np_ != op:
            if np_ < op:
                np_ *= new_shape[nj]; nj += 1
            else:
                op *= oshape[oj]; oj += 1
        for ok in range(oi, oj - 1):
            if ostride[ok] != oshape[ok + 1] * ostride[ok + 1]:
                return False
        new_strides[nj - 1] = ostride[oj - 1]
        for ok in range(nj - 1, ni, -1):
            new_strides[ok - 1] = new_strides[ok] * new_shape[ok]
        ni = nj; nj += 1
        oi = oj; oj += 1
    cdef int64_t last = new_strides[ni - 1] if ni > 0 else 1
    for i in range(ni, new_nd):
        new_strides[i] = last
    return True


def reshape(a, shape, order='C'):
    """np.reshape (intern slot _RESHAPE; forward.pyx:180-186): a view when the
    strides allow it, otherwise a compacting copy."""
    cdef ndarray s = _as_device(a)
    cdef int64_t shp[8]
    cdef int64_t st[8]
    cdef int nd, i, unknown = -1
    cdef int64_t known = 1, total = s._numel()
    if not isinstance(shape, (tuple, list)):
        shape = (shape,)
    nd = len(shape)
    if nd > SK_MAX_NDIM:
        raise ValueError('too many dimensions')
    for i in range(nd):
        shp[i] = shape[i]
        if shp[i] == -1:
            if unknown >= 0:
                raise ValueError('can only specify one unknown dimension')
            unknown = i
        else:
            known *= shp[i]
    if unknown >= 0:
        if known == 0 or total % known != 0:
            raise ValueError(f'cannot reshape array of size {total} into shape {tuple(shape)}')
        shp[unknown] = total // known
    elif known != total:
        raise ValueError(f'cannot reshape array of size {total} into shape {tuple(shape)}')
    cdef int64_t n = 1
    if total == 0 or s._is_contiguous():
        for i in range(nd - 1, -1, -1):
            st[i] = n
            n *= shp[i]
        return s._view(nd, shp, st, 0)
    if _nocopy_reshape(s, nd, shp, st):
        return s._view(nd, shp, st, 0)
    cdef ndarray c = s._compact()
    for i in range(nd - 1, -1, -1):
        st[i] = n
        n *= shp[i]
    return c._view(nd, shp, st, 0)


def squeeze(a, axis=None):
    """np.squeeze (intern slot _SQUEEZE; forward.pyx:244-245): a view."""
    cdef ndarray s = _as_device(a)
    cdef int64_t shp[8]
    cdef int64_t st[8]
    cdef int nd = 0, i, ax
    cdef unsigned int drop = 0
    if axis is not None:
        if not isinstance(axis, (tuple, list)):
            axis = (axis,)
        for ax in axis:
            if ax < 0:
                ax += s._ndim
            if ax < 0 or ax >= s._ndim:
                raise ValueError('axis out of bounds')
            if s._shape[ax] != 1:
                raise ValueError('cannot select an axis to squeeze out which has size not equal to one')
            drop |= (1u << ax)
    for i in range(s._ndim):
        if (axis is None and s._shape[i] == 1) or (drop & (1u << i)):
            continue
        shp[nd] = s._shape[i]
        st[nd] = s._strides[i]
        nd += 1
    return s._view(nd, shp, st, 0)


def expand_dims(a, axis):
    cdef ndarray s = _as_device(a)
    cdef int64_t shp[8]
    cdef int64_t st[8]
    cdef int i, j = 0, ax = axis, nd = s._ndim + 1
    if ax < 0:
        ax += nd
    for i in range(nd):
        if i == ax:
            shp[i] = 1; st[i] = 0
        else:
            shp[i] = s._shape[j]; st[i] = s._strides[j]; j += 1
    return s._view(nd, shp, st, 0)


def ascontiguousarray(a):
    return _as_device(a)._compact()


# --------------------------------------------------------------------------- indexing
cdef object _getitem(ndarray a, object idx):
    """Basic slicing -> view; a single integer array (first axis) -> row gather."""
    cdef int64_t shp[8]
    cdef int64_t st[8]
    cdef int64_t offset = 0, i64, dim
    cdef int nd = 0, axis = 0, n_specified = 0, n_fill, k
    cdef Py_ssize_t start, stop, step, length
    if type(idx) is ndarray or isinstance(idx, (list, _np_ndarray)):
        return _gather(a, idx)
    if not isinstance(idx, tuple):
        idx = (idx,)
    for it in idx:
        if it is not None and it is not Ellipsis:
            n_specified += 1
        if type(it) is ndarray or isinstance(it, (list, _np_ndarray)):
            if len(idx) >= 1 and idx[0] is it:
                for other in idx[1:]:
                    if not (isinstance(other, slice) and other == slice(None)):
                        raise IndexError('soket_b200: only x[int_array] / x[int_array, :, ...] advanced indexing is supported')
                return _gather(a, it)
            raise IndexError('soket_b200: only x[int_array] / x[int_array, :, ...] advanced indexing is supported')
    if n_specified > a._ndim:
        raise IndexError(f'too many indices for array: array is {a._ndim}-dimensional, but {n_specified} were indexed')
    for it in idx:
        if it is Ellipsis:
            n_fill = a._ndim - n_specified
            for k in range(n_fill):
                shp[nd] = a._shape[axis]; st[nd] = a._strides[axis]
                nd += 1; axis += 1
            n_specified = a._ndim  # a second Ellipsis expands to nothing
        elif it is None:
            shp[nd] = 1; st[nd] = 0; nd += 1
        elif isinstance(it, slice):
            dim = a._shape[axis]
            start, stop, step = it.indices(dim)
            length = len(range(start, stop, step))
            offset += start * a._strides[axis]
            shp[nd] = length
            st[nd] = a._strides[axis] * step
            nd += 1; axis += 1
        else:
            try:
                i64 = it.__index__()
            except AttributeError:
                raise IndexError('only integers, slices (`:`), ellipsis (`...`), None and integer arrays are valid indices')
            dim = a._shape[axis]
            if i64 < -dim or i64 >= dim:
                raise IndexError(f'index {i64} is out of bounds for axis {axis} with size {dim}')
            if i64 < 0:
                i64 += dim
            offset += i64 * a._strides[axis]
            axis += 1
    while axis < a._ndim:
        shp[nd] = a._shape[axis]; st[nd] = a._strides[axis]
        nd += 1; axis += 1
    if nd > SK_MAX_NDIM:
        raise IndexError('too many dimensions')
    return a._view(nd, shp, st, offset)


cdef ndarray _gather(ndarray a, object index):
    cdef ndarray ix
    if type(index) is ndarray:
        ix = <ndarray> index
    else:
        host = np.asarray(index)
        if host.dtype == np.bool_:
            raise IndexError('soket_b200: boolean mask indexing is not supported')
        if host.dtype.kind not in 'iu':
            raise IndexError('arrays used as indices must be of integer type')
        ix = _from_numpy(host, _code(host.dtype))
    if ix._code == SK_BOOL or ix._code >= SK_F16:
        raise IndexError('arrays used as indices must be of integer type')
    if a._ndim == 0:
        raise IndexError('too many indices for array')
    # rows of `a` must be contiguous for the gather kernel
    cdef ndarray src = a
    cdef int64_t inner = 1
    cdef int i
    for i in range(a._ndim - 1, 0, -1):
        if a._shape[i] != 1 and a._strides[i] != inner:
            src = a._compact()
            break
        inner *= a._shape[i]
    cdef int64_t shp[8]
    cdef int nd = 0
    if ix._ndim + src._ndim - 1 > SK_MAX_NDIM:
        raise IndexError('too many dimensions')
    for i in range(ix._ndim):
        shp[nd] = ix._shape[i]; nd += 1
    for i in range(1, src._ndim):
        shp[nd] = src._shape[i]; nd += 1
    cdef ndarray out = _new_array(nd, shp, src._code)
    cdef sk_array s, x, o
    src._desc(&s); ix._desc(&x); out._desc(&o)
    _check(sk_gather_rows(&s, &x, &o))
    return out


cdef int _setitem(ndarray a, object idx, object value) except -1:
    if a._readonly:
        raise ValueError('assignment destination is read-only')
    cdef object target = _getitem(a, idx)
    cdef ndarray view = <ndarray> target
    cdef ndarray src
    cdef sk_array s, d
    a._touch()
    if isinstance(value, (bool, int, float)):
        _fill(view, value)
        return 0
    src = _as_device(value) if not isinstance(value, (list, tuple)) else array(value, view._np_dtype)
    # NumPy allows extra leading 1-dims on the value
    while src._ndim > view._ndim and src._shape[0] == 1:
        src = src._view(src._ndim - 1, &src._shape[1], &src._strides[1], 0)
    view._desc(&d)
    src._desc_bcast(&s, view._ndim, view._shape)
    _check(sk_copy(&s, &d))
    return 0


def stack(arrays, axis=0, out=None):
    """np.stack (intern slot _STACK; soket/tensor/util.pyx:36)."""
    arrays = [_as_device(x) for x in arrays]
    if not arrays:
        raise ValueError('need at least one array to stack')
    cdef ndarray first = <ndarray> arrays[0]
    cdef ndarray cur, slot
    cdef int nd = first._ndim + 1, ax = axis, i, j, k
    cdef int code = first._code
    cdef int64_t shp[8]
    cdef int64_t vshape[8]
    cdef int64_t vstr[8]
    cdef sk_array s, d
    if ax < 0:
        ax += nd
    if ax < 0 or ax >= nd:
        raise ValueError('axis out of bounds')
    for x in arrays[1:]:
        cur = <ndarray> x
        if cur.shape != first.shape:
            raise ValueError('all input arrays must have the same shape')
        code = _code(np.result_type(_NP_OF[code], cur._np_dtype))
    j = 0
    for i in range(nd):
        if i == ax:
            shp[i] = len(arrays)
        else:
            shp[i] = first._shape[j]; j += 1
    cdef ndarray res = _new_array(nd, shp, code)
    for k in range(len(arrays)):
        cur = <ndarray> arrays[k]
        j = 0
        for i in range(nd):
            if i == ax:
                continue
            vshape[j] = res._shape[i]; vstr[j] = res._strides[i]; j += 1
        slot = res._view(nd - 1, vshape, vstr, k * res._strides[ax])
        cur._desc(&s); slot._desc(&d)
        _check(sk_copy(&s, &d))
    return res


# --------------------------------------------------------------------------- elementwise
cdef inline bint _is_pyscalar(object x):
    return type(x) is float or type(x) is int or type(x) is bool


cdef object _bcast_shape(ndarray a, ndarray b, int64_t *shp, int *nd_out):
    cdef int nd = a._ndim if a._ndim > b._ndim else b._ndim
    cdef int i, ia, ib
    cdef int64_t da, db
    for i in range(nd):
        ia = i - (nd - a._ndim)
        ib = i - (nd - b._ndim)
        da = a._shape[ia] if ia >= 0 else 1
        db = b._shape[ib] if ib >= 0 else 1
        if da == db or db == 1:
            shp[i] = da
        elif da == 1:
            shp[i] = db
        else:
            raise ValueError(f'operands could not be broadcast together with shapes {a.shape} {b.shape}')
    nd_out[0] = nd
    return None


cdef object _binary(int op, object x, object y, object dtype):
    """NumPy-semantics binary ufunc: broadcasting, NEP-50 weak Python scalars,
    `dtype=` selects the computation / output type."""
    cdef ndarray a, b, out
    cdef sk_array da, db, dout
    cdef int64_t shp[8]
    cdef int nd = 0, code
    cdef bint is_cmp = op >= SK_OP_EQ
    cdef bint xs = _is_pyscalar(x), ys = _is_pyscalar(y)
    cdef object rdt
    if op == SK_OP_DIV and dtype is not None and _code(dtype) < SK_F16:
        # numpy.divide has float loops only: a non-float `dtype=` is rejected (this is what a Soket
        # true-divide of two integer tensors hits on the reference's CPU backend, forward.pyx:75,85)
        raise TypeError('No loop matching the specified signature and casting was found for ufunc divide')
    if xs and ys:
        x = np.asarray(x)
        xs = False
    if xs or ys:
        a = _as_device(y if xs else x)
        scalar = x if xs else y
        if is_cmp:
            code = SK_BOOL
        elif dtype is not None:
            code = _code(dtype)
        else:
            rdt = np.result_type(a._np_dtype, scalar)
            if op == SK_OP_DIV and rdt.kind in 'iub':
                rdt = _F64
            code = _code(rdt)
        if op == SK_OP_SUB and code == SK_BOOL:
            raise TypeError('numpy boolean subtract, the `-` operator, is not supported, use the bitwise_xor, '
                            'the `^` operator, or the logical_xor function instead.')
        out = _new_array(a._ndim, a._shape, code)
        a._desc(&da)
        out._desc(&dout)
        if type(scalar) is float:
            _check(sk_ewise_scalar(op, &da, <double> scalar, 0, 0, 1 if xs else 0, &dout))
        else:
            _check(sk_ewise_scalar(op, &da, 0.0, <int64_t> int(scalar), 1, 1 if xs else 0, &dout))
        return out
    a = _as_device(x)
    b = _as_device(y)
    if is_cmp:
        code = SK_BOOL
    elif dtype is not None:
        code = _code(dtype)
    elif a._code == b._code and not (op == SK_OP_DIV and a._code < SK_F16):
        code = a._code
    else:
        rdt = np.result_type(a._np_dtype, b._np_dtype)
        if op == SK_OP_DIV and rdt.kind in 'iub':
            rdt = _F64
        code = _code(rdt)
    if op == SK_OP_SUB and code == SK_BOOL:
        raise TypeError('numpy boolean subtract, the `-` operator, is not supported, use the bitwise_xor, '
                        'the `^` operator, or the logical_xor function instead.')
    _bcast_shape(a, b, shp, &nd)
    out = _new_array(nd, shp, code)
    a._desc_bcast(&da, nd, shp)
    b._desc_bcast(&db, nd, shp)
    out._desc(&dout)
    _check(sk_ewise_binary(op, &da, &db, &dout))
    return out


def add(x, y, out=None, dtype=None): return _binary(SK_OP_ADD, x, y, dtype)
def subtract(x, y, out=None, dtype=None): return _binary(SK_OP_SUB, x, y, dtype)
def multiply(x, y, out=None, dtype=None): return _binary(SK_OP_MUL, x, y, dtype)
def divide(x, y, out=None, dtype=None): return _binary(SK_OP_DIV, x, y, dtype)
def power(x, y, out=None, dtype=None): return _binary(SK_OP_POW, x, y, dtype)
def maximum(x, y, out=None, dtype=None): return _binary(SK_OP_MAXIMUM, x, y, dtype)
def minimum(x, y, out=None, dtype=None): return _binary(SK_OP_MINIMUM, x, y, dtype)
def equal(x, y, out=None): return _binary(SK_OP_EQ, x, y, None)
def not_equal(x, y, out=None): return _binary(SK_OP_NE, x, y, None)
def greater(x, y, out=None): return _binary(SK_OP_GT, x, y, None)
def greater_equal(x, y, out=None): return _binary(SK_OP_GE, x, y, None)
def less(x, y, out=None): return _binary(SK_OP_LT, x, y, None)
def less_equal(x, y, out=None): return _binary(SK_OP_LE, x, y, None)
true_divide = divide


cdef object _unary(int op, object x, bint float_result):
    cdef ndarray a = _as_device(np.asarray(x)) if _is_pyscalar(x) else _as_device(x)
    cdef int code = a._code
    if float_result and code < SK_F16:
        # NumPy: exp/log of integers computes in float64 (float16 for 8-bit, float32 for 16-bit)
        code = _code(np.result_type(a._np_dtype, np.float16))
    if op == SK_UOP_NEG and code == SK_BOOL:
        raise TypeError('The numpy boolean negative, the `-` operator, is not supported')
    cdef ndarray out = _new_array(a._ndim, a._shape, code)
    cdef sk_array da, dout
    a._desc(&da)
    out._desc(&dout)
    _check(sk_ewise_unary(op, &da, &dout))
    return out


def negative(x, out=None): return _unary(SK_UOP_NEG, x, False)
def exp(x, out=None): return _unary(SK_UOP_EXP, x, True)
def log(x, out=None): return _unary(SK_UOP_LOG, x, True)
def sqrt(x, out=None): return _unary(SK_UOP_SQRT, x, True)
def absolute(x, out=None): return _unary(SK_UOP_ABS, x, False)
abs = absolute


def relu_backward(x, adj):
    """(x > 0) * adj in one pass (backward.pyx:849-874)."""
    cdef ndarray a = _as_device(x)._compact()
    cdef ndarray g = _as_device(adj)._compact()
    cdef ndarray out = _new_array(a._ndim, a._shape, SK_F32)
    cdef sk_array da, dg, dout
    a._desc(&da); g._desc(&dg); out._desc(&dout)
    _check(sk_relu_bwd(&da, &dg, &dout))
    return out


# --------------------------------------------------------------------------- reductions
cdef uint32_t _axes_mask(object axis, int ndim) except? 0xFFFFFFFF:
    cdef uint32_t mask = 0
    cdef int ax
    if axis is None:
        return (1u << ndim) - 1 if ndim < 32 else 0xFFFFFFFFu
    if not isinstance(axis, (tuple, list)):
        axis = (axis,)
    for v in axis:
        ax = v
        if ax < 0:
            ax += ndim
        if ax < 0 or ax >= ndim:
            raise ValueError(f'axis {v} is out of bounds for array of dimension {ndim}')
        if mask & (1u << ax):
            raise ValueError('duplicate value in axis')
        mask |= (1u << ax)
    return mask


cdef object _reduce(int op, object x, object axis, object dtype, bint keepdims):
    cdef ndarray a = _as_device(x)
    cdef uint32_t mask = _axes_mask(axis, a._ndim)
    cdef int code, i, nd = 0, knd = 0
    cdef int64_t shp[8]
    cdef int64_t kshp[8]
    cdef int64_t kst[8]
    cdef sk_array da, dout
    if dtype is not None:
        code = _code(dtype)
    elif op == SK_RED_SUM:
        # NumPy: integer sums accumulate in (u)int64, bool in int64
        if a._code == SK_BOOL or a._code in (SK_I8, SK_I16, SK_I32):
            code = SK_I64
        elif a._code in (SK_U8, SK_U16, SK_U32):
            code = SK_U64
        else:
            code = a._code
    elif op == SK_RED_MEAN:
        code = SK_F64 if a._code < SK_F16 else a._code
    else:
        code = a._code
    for i in range(a._ndim):
        if mask & (1u << i):
            kshp[knd] = 1; knd += 1
        else:
            shp[nd] = a._shape[i]; nd += 1
            kshp[knd] = a._shape[i]; knd += 1
    cdef ndarray out = _new_array(nd, shp, code)
    a._desc(&da)
    out._desc(&dout)
    _check(sk_reduce(op, &da, mask, &dout))
    if keepdims:
        nd = 1
        for i in range(knd - 1, -1, -1):
            kst[i] = nd
            nd *= kshp[i]
        return out._view(knd, kshp, kst, 0)
    return out


def sum(a, axis=None, dtype=None, out=None, keepdims=False):
    """np.sum(x, axes, dtype, out, keepdims) (intern slot _SUM; forward.pyx:128-137)."""
    return _reduce(SK_RED_SUM, a, axis, dtype, keepdims is True)


def mean(a, axis=None, dtype=None, out=None, keepdims=False):
    """np.mean(x, axes, dtype, out, keepdims) (intern slot _MEAN; forward.pyx:140-149)."""
    return _reduce(SK_RED_MEAN, a, axis, dtype, keepdims is True)


def max(a, axis=None, out=None, keepdims=False):
    """np.max(x, axes, out, keepdims) (intern slot _MAX; forward.pyx:152-160)."""
    return _reduce(SK_RED_MAX, a, axis, None, keepdims is True)


def min(a, axis=None, out=None, keepdims=False):
    """np.min(x, axes, out, keepdims) (intern slot _MIN; forward.pyx:162-170)."""
    return _reduce(SK_RED_MIN, a, axis, None, keepdims is True)


amax = max
amin = min


cdef object _argreduce(int is_min, object x, object axis, bint keepdims):
    cdef ndarray a = _as_device(x)
    cdef int ax = -1, i, nd = 0, knd = 0
    cdef int64_t shp[8]
    cdef int64_t kshp[8]
    cdef int64_t kst[8]
    cdef sk_array da, dout
    if axis is None:
        a = a._compact()
        for i in range(a._ndim):
            kshp[knd] = 1; knd += 1
    else:
        ax = axis
        if ax < 0:
            ax += a._ndim
        if ax < 0 or ax >= a._ndim:
            raise ValueError(f'axis {axis} is out of bounds for array of dimension {a._ndim}')
        for i in range(a._ndim):
            if i == ax:
                kshp[knd] = 1; knd += 1
            else:
                shp[nd] = a._shape[i]; nd += 1
                kshp[knd] = a._shape[i]; knd += 1
    cdef ndarray out = _new_array(nd, shp, SK_I64)
    a._desc(&da)
    out._desc(&dout)
    _check(sk_argreduce(is_min, &da, ax, &dout))
    if keepdims:
        nd = 1
        for i in range(knd - 1, -1, -1):
            kst[i] = nd
            nd *= kshp[i]
        return out._view(knd, kshp, kst, 0)
    return out


def argmax(a, axis=None, out=None, *, keepdims=False):
    """np.argmax(x, axis, keepdims=) -> int64 (intern slot _ARGMAX; tensor.pyx:852-859)."""
    return _argreduce(0, a, axis, keepdims is True)


def argmin(a, axis=None, out=None, *, keepdims=False):
    return _argreduce(1, a, axis, keepdims is True)


# --------------------------------------------------------------------------- matmul
cdef int _DEFAULT_MM_ALGO = SK_MM_AUTO

MM_AUTO = SK_MM_AUTO
MM_SIMT = SK_MM_SIMT
MM_TF32X3 = SK_MM_TF32X3
MM_TF32 = SK_MM_TF32
MM_BF16 = SK_MM_BF16
MM_F16X3 = SK_MM_F16X3


def set_matmul_algo(int algo):
    """Select the fp32 matmul path globally (MM_AUTO / MM_SIMT / MM_TF32X3 / MM_TF32)."""
    global _DEFAULT_MM_ALGO
    _DEFAULT_MM_ALGO = algo


cdef ndarray _mm_operand(ndarray a):
    """Operand the GEMM kernels can consume in place: the last two axes are
    row-major or transposed (one unit stride); anything else is compacted."""
    cdef int nd = a._ndim
    cdef int64_t s0 = a._strides[nd - 2], s1 = a._strides[nd - 1]
    cdef int64_t d0 = a._shape[nd - 2], d1 = a._shape[nd - 1]
    if (s1 == 1 or d1 == 1) and (s0 >= d1 or d0 == 1):
        return a
    if (s0 == 1 or d0 == 1) and (s1 >= d0 or d1 == 1):
        return a
    return a._compact()


cdef object _matmul_impl(object x, object y, object bias, int epilogue, int algo, object dtype=None):
    cdef ndarray a = _as_device(x)
    cdef ndarray b = _as_device(y)
    cdef ndarray bi = None
    cdef ndarray out
    cdef sk_array da, db, dbias, dout
    cdef int64_t shp[8]
    cdef int nd, i, ia, ib, nba, nbb, nb
    cdef int64_t xa, xb
    cdef bint a_vec = a._ndim == 1, b_vec = b._ndim == 1
    if a._ndim == 0 or b._ndim == 0:
        raise ValueError('matmul: input operand does not have enough dimensions')
    if a_vec:
        a = a._view(2, [1, a._shape[0]], [0, a._strides[0]], 0)
    if b_vec:
        b = b._view(2, [b._shape[0], 1], [b._strides[0], 0], 0)
    if a._shape[a._ndim - 1] != b._shape[b._ndim - 2]:
        raise ValueError(f'matmul: input operand 1 has a mismatch in its core dimension 0 '
                         f'(size {b._shape[b._ndim - 2]} is different from {a._shape[a._ndim - 1]})')
    # result dtype = the `dtype=` keyword (forward.pyx:172-178 passes the promoted Tensor dtype) or
    # NumPy's promotion of the operands; float32 / float16 results are computed by the fp32 GEMMs,
    # float64 and integer results by the float64 kernel (exact for integers below 2^53)
    cdef int rcode, ccode
    if dtype is not None:
        rcode = _code(dtype)
    elif a._code == SK_BF16 and b._code == SK_BF16:
        rcode = SK_F32
    else:
        rcode = _code(np.result_type(a._np_dtype, b._np_dtype))
    if a._code == SK_BF16 and b._code == SK_BF16 and rcode == SK_F32:
        ccode = SK_BF16
    else:
        ccode = SK_F32 if (rcode == SK_F32 or rcode == SK_F16) else SK_F64
        if ccode == SK_F64 and (bias is not None or epilogue != SK_EPI_NONE):
            raise TypeError('linear: the fused epilogues are float32 only')
        a = _cast_copy(a, ccode) if a._code != ccode else a
        b = _cast_copy(b, ccode) if b._code != ccode else b
    a = _mm_operand(a)
    b = _mm_operand(b)
    nba = a._ndim - 2; nbb = b._ndim - 2
    nb = nba if nba > nbb else nbb
    for i in range(nb):
        ia = i - (nb - nba); ib = i - (nb - nbb)
        xa = a._shape[ia] if ia >= 0 else 1
        xb = b._shape[ib] if ib >= 0 else 1
        if xa != xb and xa != 1 and xb != 1:
            raise ValueError('matmul: batch dimensions do not broadcast')
        shp[i] = xa if xa != 1 else xb
    shp[nb] = a._shape[a._ndim - 2]
    shp[nb + 1] = b._shape[b._ndim - 1]
    nd = nb + 2
    out = _new_array(nd, shp, SK_F64 if ccode == SK_F64 else SK_F32)
    a._desc(&da); b._desc(&db); out._desc(&dout)
    if a._code == SK_BF16:
        algo = SK_MM_BF16
    if bias is not None:
        bi = _as_device(bias)._compact()
        bi._desc(&dbias)
        _check(sk_linear_fwd(&da, &db, &dbias, &dout, epilogue, algo))
    elif epilogue != SK_EPI_NONE:
        _check(sk_linear_fwd(&da, &db, NULL, &dout, epilogue, algo))
    else:
        _check(sk_matmul(&da, &db, &dout, algo))
    if out._code != rcode:
        out = _cast_copy(out, rcode)
    if a_vec and b_vec:
        return out._view(0, shp, shp, 0)
    if a_vec:
        return squeeze(out, nd - 2)
    if b_vec:
        return squeeze(out, nd - 1)
    return out


def matmul(x, y, out=None, dtype=None, algo=None):
    """np.matmul(x, y, dtype='float32') (intern slot _MATMUL; forward.pyx:172-178,
    backward.pyx:720-736).  `.T` views are consumed in place."""
    return _matmul_impl(x, y, None, SK_EPI_NONE, _DEFAULT_MM_ALGO if algo is None else <int> algo, dtype)


def linear(x, w, bias=None, relu=False, algo=None):
    """Fused relu?(x @ w + bias)  (prototypes.pyx:108-115 + :302)."""
    cdef int epi
    if bias is not None:
        epi = SK_EPI_BIAS_RELU if relu else SK_EPI_BIAS
    else:
        epi = SK_EPI_RELU if relu else SK_EPI_NONE
    return _matmul_impl(x, w, bias, epi, _DEFAULT_MM_ALGO if algo is None else <int> algo)


def linear_bwd(adj, x, w, bint want_bias=False, out_dw=None, out_db=None):
    """Backward of y = x @ w (backward.pyx:704-742): returns (adj @ w.T, x.T @ adj) from one
    call, so that the fp16x3 path splits `adj` once for both GEMMs.  want_bias: also the bias
    gradient adj.sum(0) (autodiff.pyx:84) from the same pass over adj -> (dx, dw, db)."""
    cdef ndarray a = _as_device(adj)._compact()
    cdef ndarray xx = _as_device(x)._compact()
    cdef ndarray ww = _as_device(w)._compact()
    if a._ndim != 2 or xx._ndim != 2 or ww._ndim != 2:
        raise ValueError('linear_bwd: operands must be 2-D')
    if a._code != SK_F32 or xx._code != SK_F32 or ww._code != SK_F32:
        raise TypeError('linear_bwd: operands must be float32')
    if xx._shape[0] != a._shape[0] or ww._shape[0] != xx._shape[1] or ww._shape[1] != a._shape[1]:
        raise ValueError('linear_bwd: shapes must be adj (B,O), x (B,I), w (I,O)')
    cdef int64_t shp[2]
    shp[0] = xx._shape[0]; shp[1] = xx._shape[1]
    cdef ndarray dx = _new_array(2, shp, SK_F32)
    shp[0] = ww._shape[0]; shp[1] = ww._shape[1]
    # out_dw / out_db: caller-owned result buffers (data-parallel training points them at the slots of
    # its flat gradient arena, so the all-reduce needs no gather copy)
    cdef ndarray dw
    if out_dw is not None:
        dw = <ndarray> out_dw
        dw._touch()
        if dw._code != SK_F32 or dw._ndim != 2 or dw._shape[0] != shp[0] or dw._shape[1] != shp[1] or not dw._is_contiguous():
            raise ValueError('linear_bwd: out_dw must be a contiguous float32 array of w\'s shape')
    else:
        dw = _new_array(2, shp, SK_F32)
    cdef sk_array da, dxx, dww, ddx, ddw
    a._desc(&da); xx._desc(&dxx); ww._desc(&dww); dx._desc(&ddx); dw._desc(&ddw)
    cdef ndarray db
    if want_bias:
        shp[0] = a._shape[1]
        if out_db is not None:
            db = <ndarray> out_db
            db._touch()
            if db._code != SK_F32 or db._numel() != shp[0] or not db._is_contiguous():
                raise ValueError('linear_bwd: out_db must be a contiguous float32 vector of length O')
        else:
            db = _new_array(1, shp, SK_F32)
        _check(sk_linear_bwd_bias(&da, &dxx, &dww, &ddx, &ddw, <float *> db._ptr))
        return dx, dw, db
    _check(sk_linear_bwd(&da, &dxx, &dww, &ddx, &ddw))
    return dx, dw


# --------------------------------------------------------------------------- pre-split GEMM operands
# (sk_split_f16 / sk_gemm_f16x3: include/soket_b200.h, "Pre-split operands of the fp16x3 GEMM")
cdef long _GRAPH_EPOCH = 0      # bumped by every graph replay (buffers rewritten behind the host's back)
cdef long _CAPTURE_ID = 0       # bumped by every Graph.begin()


cdef long _graph_epoch():
    return _GRAPH_EPOCH


cdef inline long _capture_now():
    """0 in eager mode, else the id of the running capture.  Derived data made DURING a capture exists
    only inside that graph (its kernels have not run): it may feed later nodes of the same capture,
    never eager code; data made eagerly may feed a capture (its buffers are real)."""
    return _CAPTURE_ID if sk_graph_capturing() else 0


cdef class SplitMat:
    cdef bint valid_for(self, ndarray x):
        """Still describes the CURRENT contents of x?  (No in-place write through any view since the
        split, no graph replay, same storage and shape; never during a capture, where the kernels
        that would have filled it have not run.)"""
        return (self.src_ptr == x._ptr and self.version == x._buf.version and self.epoch == _GRAPH_EPOCH
                and x._ndim == 2 and x._shape[0] == self.rows and x._shape[1] == self.cols
                and (self.capture == 0 or self.capture == _capture_now()))

    def __repr__(self):
        return f'SplitMat({self.rows} x {self.cols}, ld {self.ld})'


cdef class AbsMax:
    pass


cdef SplitMat _new_split(int64_t rows, int64_t cols):
    cdef SplitMat m = SplitMat.__new__(SplitMat)
    cdef int64_t shp[2]
    m.rows = rows; m.cols = cols
    m.ld = (cols + 7) // 8 * 8
    shp[0] = rows; shp[1] = m.ld
    m.hi = _new_array(2, shp, SK_F16)
    m.lo = _new_array(2, shp, SK_F16)
    shp[0] = 4
    m.scale = _new_array(1, shp, SK_F32)
    m.src_ptr = 0; m.version = -1; m.epoch = -1; m.capture = 0
    m.amax = None
    return m


cdef void _bind_split(SplitMat m, ndarray x):
    """m holds the split of x's current contents: remember that on x (until x is written to)."""
    m.src_ptr = x._ptr
    m.version = x._buf.version
    m.epoch = _GRAPH_EPOCH
    m.capture = _capture_now()
    x._meta = m


def new_absmax_word(ndarray owner=None):
    """A zeroed device word for a kernel to atomicMax |x| bit patterns into (sk_ln_extras.dx_absmax)."""
    cdef AbsMax a = AbsMax.__new__(AbsMax)
    cdef int64_t one = 1
    a.word = _new_array(1, &one, SK_U32)
    _check(sk_memset(<void *> a.word._ptr, 0, 4))
    a.version = -1; a.epoch = -1; a.capture = 0
    a.rotating = False
    return a


cdef AbsMax _new_rotating_absmax():
    """{max |x| (0 = not known yet), accumulator}: the words sk_adam_step_split keeps current by itself."""
    cdef AbsMax a = AbsMax.__new__(AbsMax)
    cdef int64_t two = 2
    a.word = _new_array(1, &two, SK_U32)
    _check(sk_memset(<void *> a.word._ptr, 0, 8))
    a.version = -1; a.epoch = -1; a.capture = 0
    a.rotating = True
    return a


def new_rotating_absmax():
    return _new_rotating_absmax()


def has_absmax(ndarray x):
    """Does x carry a |max| word that is still valid for its contents (so that split_f16 needs no
    temporary of its own -- a requirement for running it off the compute stream)?"""
    cdef AbsMax am
    if not isinstance(x._meta, AbsMax):
        return False
    am = <AbsMax> x._meta
    return (am.version == x._buf.version and am.epoch == _GRAPH_EPOCH
            and (am.capture == 0 or am.capture == _capture_now()))


def bind_absmax_word(ndarray word, ndarray x):
    """Attach a device word a kernel has filled with the bit pattern of max |x| to x."""
    cdef AbsMax a = AbsMax.__new__(AbsMax)
    a.word = word
    a.rotating = False
    bind_absmax(a, x)


def bind_absmax(AbsMax a, ndarray x):
    a.version = x._buf.version
    a.epoch = _GRAPH_EPOCH
    a.capture = _capture_now()
    x._meta = a


def split_f16(x, want_colsum=False, out_colsum=None, out_hi=None, out_lo=None):
    """float32 (rows, cols) -> SplitMat (and the column sums of x when asked: the bias gradient rides
    along with the adjoint's split, autodiff.pyx:84).  Uses the |max| a producing kernel attached to
    x (AbsMax) when there is one, else computes it in an extra pass.  out_hi / out_lo: caller-owned
    contiguous float16 (rows, cols) arrays to split into (cols % 8 == 0)."""
    cdef ndarray a = _as_device(x)
    if a._ndim != 2 or a._code != SK_F32:
        raise TypeError('split_f16: expected a 2-D float32 array')
    a = a._compact()
    if a._shape[1] % 4 != 0 and a._shape[0] > 1:
        raise ValueError('split_f16: the row length must be a multiple of 4 elements (16-byte aligned rows)')
    cdef SplitMat m
    cdef int64_t four = 4
    if out_hi is not None:
        m = SplitMat.__new__(SplitMat)
        m.rows = a._shape[0]; m.cols = a._shape[1]; m.ld = m.cols
        m.hi = <ndarray> out_hi; m.lo = <ndarray> out_lo
        for h in (m.hi, m.lo):
            if ((<ndarray> h)._code != SK_F16 or (<ndarray> h)._numel() != m.rows * m.cols or not (<ndarray> h)._is_contiguous()
                    or m.cols % 8 != 0):
                raise ValueError('split_f16: out_hi / out_lo must be contiguous float16 arrays of x\'s size (cols % 8 == 0)')
        m.scale = _new_array(1, &four, SK_F32)
        m.src_ptr = 0; m.version = -1; m.epoch = -1; m.capture = 0
        m.amax = None
    else:
        m = _new_split(a._shape[0], a._shape[1])
    cdef const unsigned int *amax = NULL
    cdef AbsMax am = None
    if isinstance(a._meta, AbsMax):
        am = <AbsMax> a._meta
        if (am.version == a._buf.version and am.epoch == _GRAPH_EPOCH
                and (am.capture == 0 or am.capture == _capture_now())):
            amax = <const unsigned int *> am.word._ptr
        else:
            am = None
    cdef ndarray cs = None
    cdef int64_t cols = a._shape[1]
    if want_colsum:
        if out_colsum is not None:
            cs = <ndarray> out_colsum
            if cs._code != SK_F32 or cs._numel() != cols or not cs._is_contiguous():
                raise ValueError('split_f16: out_colsum must be a contiguous float32 vector of `cols` elements')
            cs._touch()
        else:
            cs = _new_array(1, &cols, SK_F32)
    _check(sk_split_f16(<const float *> a._ptr, a._shape[0], a._shape[1], a._strides[0] if a._shape[0] > 1 else a._shape[1],
                        amax, <void *> m.hi._ptr, <void *> m.lo._ptr, m.ld, <float *> m.scale._ptr,
                        <float *> cs._ptr if cs is not None else NULL))
    _bind_split(m, a)
    if am is not None and am.rotating:
        m.amax = am            # the optimizer refreshes hi / lo itself from now on (sk_adam_step_split)
    if want_colsum:
        return m, cs
    return m


def rebind_split(SplitMat m, ndarray x):
    """m's hi / lo were rewritten (by a kernel the caller launched) to describe x's CURRENT contents."""
    _bind_split(m, x)


def get_split(x):
    """The SplitMat of x: the one a producing kernel (or an earlier call) attached to it if x has not
    been written to since, else a fresh split."""
    cdef ndarray a = <ndarray> x
    if isinstance(a._meta, SplitMat) and (<SplitMat> a._meta).valid_for(a):
        return a._meta
    return split_f16(a)


def gemm_split_supported(int64_t M, int64_t N, int64_t K):
    return bool(sk_gemm_f16x3_supported(M, N, K))


def gemm_split(SplitMat a, bint a_trans, SplitMat b, bint b_trans, bias=None, bint relu=False, out=None,
               bint accumulate=False):
    """epi(op(A) @ op(B) [+ out]) on the tcgen05 fp16x3 kernel from pre-split operands.  A is the
    matrix `a` (a.rows x a.cols) or, with a_trans, its transpose -- the SAME hi / lo arrays, consumed
    MN-major instead of K-major; likewise B.  Returns `out` (M, N) float32."""
    cdef int64_t M = a.cols if a_trans else a.rows
    cdef int64_t K = a.rows if a_trans else a.cols
    cdef int64_t Kb = b.cols if b_trans else b.rows
    cdef int64_t N = b.rows if b_trans else b.cols
    if K != Kb:
        raise ValueError(f'gemm_split: inner dimensions differ ({K} vs {Kb})')
    cdef sk_split_operand oa, ob
    oa.hi = <const void *> a.hi._ptr; oa.lo = <const void *> a.lo._ptr; oa.ld = a.ld
    oa.mn_major = 1 if a_trans else 0            # A stored (M, K): K-major; stored (K, M): MN-major
    oa.scale = <const float *> a.scale._ptr
    ob.hi = <const void *> b.hi._ptr; ob.lo = <const void *> b.lo._ptr; ob.ld = b.ld
    ob.mn_major = 0 if b_trans else 1            # B stored (K, N): MN-major; stored (N, K): K-major
    ob.scale = <const float *> b.scale._ptr
    cdef int64_t shp[2]
    cdef ndarray c
    if out is not None:
        c = <ndarray> out
        if c._code != SK_F32 or c._ndim != 2 or c._shape[0] != M or c._shape[1] != N or not c._is_contiguous():
            raise ValueError('gemm_split: out must be a contiguous float32 (M, N) array')
        c._touch()
    else:
        if accumulate:
            raise ValueError('gemm_split: accumulate needs out')
        shp[0] = M; shp[1] = N
        c = _new_array(2, shp, SK_F32)
    cdef ndarray bi = None
    cdef int epi
    if bias is not None:
        bi = _as_device(bias)._compact()
        if bi._code != SK_F32 or bi._numel() != N:
            raise ValueError('gemm_split: bias must be a float32 vector of N elements')
        epi = SK_EPI_BIAS_RELU if relu else SK_EPI_BIAS
    else:
        epi = SK_EPI_RELU if relu else SK_EPI_NONE
    _check(sk_gemm_f16x3(&oa, &ob, <float *> c._ptr, N, M, N, K, <const float *> bi._ptr if bi is not None else NULL,
                         epi, accumulate))
    return c


# --------------------------------------------------------------------------- random
class _Random:
    """`backend.random` namespace (soket/backend/device.pyx:64-66)."""

    @staticmethod
    def seed(s=None):
        _check(sk_rng_seed(<uint64_t> (0 if s is None else int(s))))

    @staticmethod
    def uniform(low=0.0, high=1.0, size=None, dtype=float):
        cdef ndarray out = _new_from_tuple(() if size is None else size, _code(dtype))
        cdef sk_array d
        out._desc(&d)
        _check(sk_rng_uniform(&d, <double> low, <double> high))
        return out

    @staticmethod
    def normal(loc=0.0, scale=1.0, size=None, dtype=float):
        cdef ndarray out = _new_from_tuple(() if size is None else size, _code(dtype))
        cdef sk_array d
        out._desc(&d)
        _check(sk_rng_normal(&d, <double> loc, <double> scale))
        return out

    @staticmethod
    def binomial(n, p, size=None, dtype='int64'):
        if n != 1:
            raise NotImplementedError('soket_b200.random.binomial supports n == 1 (Bernoulli) only')
        cdef ndarray out = _new_from_tuple(() if size is None else size, _code(dtype))
        cdef sk_array d
        out._desc(&d)
        _check(sk_rng_bernoulli(&d, <double> p))
        return out

    @staticmethod
    def rand(*shape):
        return _Random.uniform(0.0, 1.0, shape)

    @staticmethod
    def randn(*shape):
        return _Random.normal(0.0, 1.0, shape)


random = _Random()


def one_hot(labels, int num_classes):
    """(rows,) integer labels -> (rows, classes) float32 one-hot, no eye() gather."""
    cdef ndarray l = _as_device(labels)._compact()
    cdef ndarray out = _new_from_tuple((l._shape[0], num_classes), SK_F32)
    cdef sk_array dl, do
    l._desc(&dl); out._desc(&do)
    _check(sk_one_hot(&dl, &do))
    return out


# --------------------------------------------------------------------------- runtime
class _DeviceHandle:
    """Stand-in for ``cupy.cuda.Device(id)`` (soket/backend/device.pyx:56-58):
    ``use()``, ``synchronize()``, context-manager protocol, ``__eq__``."""

    def __init__(self, device=0):
        self.id = 0 if device is None else int(device)

    def use(self):
        _check(sk_init(self.id))
        return self

    def synchronize(self):
        _check(sk_sync())

    def __enter__(self):
        self.use()
        return self

    def __exit__(self, *exc):
        return False

    def __eq__(self, other):
        return isinstance(other, _DeviceHandle) and other.id == self.id

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(('soket_b200.Device', self.id))

    def __repr__(self):
        return f'<soket_b200 Device {self.id}>'


def device_count():
    cdef int n = 0
    sk_device_count(&n)
    return n


def is_available():
    return device_count() > 0


def init(int device=0):
    _check(sk_init(device))


def synchronize():
    _check(sk_sync())


def launch_count():
    return <object> sk_launch_count()


def flush_l2():
    _check(sk_flush_l2())


PROF_FAMILIES = ('gemm_tc', 'gemm_simt', 'ln_fwd', 'ln_bwd', 'ewise', 'reduce', 'optim', 'copy', 'bn', 'loss', 'gemm_prep')


def profile_enable(bint on=True):
    """Bracket every instrumented kernel family with CUDA events (roofline numbers)."""
    _check(sk_prof_enable(on))


def profile_reset():
    _check(sk_prof_reset())


def profile_collect():
    """{family: {'launches', 'ms', 'work'}}; work = flops (gemm_*) or algorithmic bytes."""
    cdef int64_t n = 0
    cdef double ms = 0, work = 0
    out = {}
    for i, name in enumerate(PROF_FAMILIES):
        _check(sk_prof_collect(i, &n, &ms, &work))
        if n:
            out[name] = {'launches': int(n), 'ms': ms, 'work': work}
    return out


def memory_stats():
    cdef size_t a = 0, b = 0, c = 0
    sk_mem_stats(&a, &b, &c)
    return {'in_use': a, 'reserved': b, 'peak_in_use': c}


def empty_cache():
    _check(sk_empty_cache())


def version():
    return sk_version().decode()


cdef class Event:
    """CUDA event on the library's compute stream (bench timing)."""
    cdef void *_ev

    def __cinit__(self):
        self._ev = NULL
        _check(sk_event_create(&self._ev))

    def __dealloc__(self):
        if self._ev != NULL:
            sk_event_destroy(self._ev)

    def record(self, int stream=0):
        """Record on the compute stream (default) or on STREAM_COMM / STREAM_COPY / STREAM_OPT."""
        if stream == 0:
            _check(sk_event_record(self._ev))
        else:
            _check(sk_event_record_on(self._ev, stream))
        return self

    def wait(self, int stream=0):
        """Make `stream` wait (on the device) for this event."""
        _check(sk_stream_wait_event(stream, self._ev))

    def synchronize(self):
        _check(sk_event_sync(self._ev))

    def elapsed_ms(self, Event end):
        cdef float ms = 0
        _check(sk_event_elapsed_ms(self._ev, end._ev, &ms))
        return ms


STREAM_COMPUTE, STREAM_COMM, STREAM_COPY, STREAM_OPT = 0, 1, 2, 3


def launch_stream(int stream=0):
    """Kernels launched by later calls go to this stream (host-side switch; see sk_launch_stream)."""
    _check(sk_launch_stream(stream))


def arena_create():
    """A private allocation arena (see sk_arena_* in include/soket_b200.h)."""
    cdef int a = 0
    _check(sk_arena_create(&a))
    return a


def arena_begin(int arena):
    _check(sk_arena_begin(arena))


def arena_end():
    _check(sk_arena_end())


def arena_destroy(int arena):
    _check(sk_arena_destroy(arena))


def prefetch_wait():
    """The compute stream waits for the most recent PinnedBuffer.prefetch_to_device."""
    _check(sk_prefetch_wait())


def is_capturing():
    return sk_graph_capturing() != 0


def rng_epoch_advance():
    """Bump the device-side counter mixed into dropout seeds (first node of a captured step)."""
    _check(sk_rng_epoch_advance())


cdef class Graph:
    """Captured CUDA graph of a static-shape step (SURVEY.md section 8f-1)."""
    cdef void *_exec
    cdef list _keep

    def __cinit__(self):
        self._exec = NULL
        self._keep = []

    def __dealloc__(self):
        if self._exec != NULL:
            sk_graph_destroy(self._exec)

    def begin(self):
        global _CAPTURE_ID
        _CAPTURE_ID += 1
        _check(sk_graph_begin())

    def end(self):
        _check(sk_graph_end(&self._exec))

    def keep(self, obj):
        """Keep device buffers referenced by the graph alive."""
        self._keep.append(obj)

    def launch(self):
        global _GRAPH_EPOCH
        if self._exec == NULL:
            raise RuntimeError('graph has not been captured')
        _GRAPH_EPOCH += 1          # a replay rewrites buffers behind the host's back: drop derived caches
        _check(sk_graph_launch(self._exec))


cdef class PinnedBuffer:
    """Page-locked host staging buffer for asynchronous H2D / D2H copies."""
    cdef void *_p
    cdef size_t _nbytes
    cdef object _arr

    def __cinit__(self, shape, dtype='float32'):
        cdef object dt = np.dtype(dtype)
        cdef object shp = tuple(shape) if isinstance(shape, (tuple, list)) else (shape,)
        cdef size_t n = 1
        for s in shp:
            n *= <size_t> s
        self._nbytes = n * dt.itemsize
        self._p = NULL
        _check(sk_host_alloc(self._nbytes, &self._p))
        cdef char[::1] mv = <char[:self._nbytes if self._nbytes else 1]> (<char *> self._p)
        self._arr = np.frombuffer(mv, dtype=dt, count=n).reshape(shp)

    def __dealloc__(self):
        self._arr = None
        if self._p != NULL:
            sk_host_free(self._p)

    @property
    def array(self):
        """NumPy view of the pinned memory."""
        return self._arr

    def copy_to_device(self, ndarray dst):
        """Async H2D on the compute stream."""
        if not dst._is_contiguous() or <size_t> dst.nbytes != self._nbytes:
            raise ValueError('PinnedBuffer.copy_to_device: size / layout mismatch')
        dst._touch()
        _check(sk_h2d_async(<void *> dst._ptr, self._p, self._nbytes))

    def prefetch_to_device(self, ndarray dst):
        """Async H2D on the copy stream: after the compute work queued so far, overlapping what
        is queued next; call `prefetch_wait()` before the first kernel that reads `dst`."""
        if not dst._is_contiguous() or <size_t> dst.nbytes != self._nbytes:
            raise ValueError('PinnedBuffer.prefetch_to_device: size / layout mismatch')
        dst._touch()
        _check(sk_h2d_prefetch(<void *> dst._ptr, self._p, self._nbytes))

    def copy_from_device(self, ndarray src):
        """Async D2H on the compute stream (synchronize before reading)."""
        if not src._is_contiguous() or <size_t> src.nbytes != self._nbytes:
            raise ValueError('PinnedBuffer.copy_from_device: size / layout mismatch')
        _check(sk_d2h_async(self._p, <const void *> src._ptr, self._nbytes))
