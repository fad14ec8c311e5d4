# _core.pxd -- the device array type shared by the Cython host modules.
from libc.stdint cimport int64_t
from soket_b200._abi cimport sk_array


cdef class Buffer:
    cdef size_t ptr
    cdef size_t nbytes


cdef class ndarray:
    cdef Buffer _buf            # owner of the allocation (shared by views)
    cdef size_t _ptr            # device address of element [0, ..., 0]
    cdef int _code              # sk_dtype
    cdef int _ndim
    cdef int64_t _shape[8]
    cdef int64_t _strides[8]    # element strides
    cdef object _np_dtype       # numpy dtype object (host metadata)
    cdef bint _readonly

    cdef int64_t _numel(self)
    cdef bint _is_contiguous(self)
    cdef void _desc(self, sk_array *d)
    cdef int _desc_bcast(self, sk_array *d, int ndim, const int64_t *shape) except -1
    cdef ndarray _view(self, int ndim, const int64_t *shape, const int64_t *strides, int64_t offset)
    cdef ndarray _compact(self)


cdef ndarray _new_array(int ndim, const int64_t *shape, int code)
cdef ndarray _as_device(object x)
cdef int _check(int rc) except -1
cdef float *_fptr(ndarray a) except NULL
