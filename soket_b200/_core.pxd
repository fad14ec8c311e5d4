# _core.pxd -- the device array type shared by the Cython host modules.
from libc.stdint cimport int64_t
from soket_b200._abi cimport sk_array


cdef class Buffer:
    cdef size_t ptr
    cdef size_t nbytes
    cdef long version           # bumped by every in-place write through any view (see ndarray._touch)
    cdef object parent          # set for a window into another Buffer's allocation (arena_view): not freed here


cdef class ndarray:
    cdef Buffer _buf            # owner of the allocation (shared by views)
    cdef size_t _ptr            # device address of element [0, ..., 0]
    cdef int _code              # sk_dtype
    cdef int _ndim
    cdef int64_t _shape[8]
    cdef int64_t _strides[8]    # element strides
    cdef object _np_dtype       # numpy dtype object (host metadata)
    cdef bint _readonly
    cdef public object _meta    # derived data valid for one Buffer.version: SplitMat | AbsMax | None

    cdef int64_t _numel(self)
    cdef bint _is_contiguous(self)
    cdef void _desc(self, sk_array *d)
    cdef int _desc_bcast(self, sk_array *d, int ndim, const int64_t *shape) except -1
    cdef ndarray _view(self, int ndim, const int64_t *shape, const int64_t *strides, int64_t offset)
    cdef ndarray _compact(self)
    cdef void _touch(self)


cdef class SplitMat:
    """A float32 matrix as the fp16x3 GEMM consumes it: X * scale = hi + lo, one scale (sk_split_f16)."""
    cdef public ndarray hi
    cdef public ndarray lo
    cdef public ndarray scale   # float32 (4,): {scale, 1/scale, |max| or bound, 0}
    cdef public int64_t rows
    cdef public int64_t cols
    cdef public int64_t ld
    cdef size_t src_ptr
    cdef long version
    cdef long epoch
    cdef long capture
    cdef public object amax     # AbsMax (rotating words) kept alive for the optimizer's fused re-split, or None
    cdef bint valid_for(self, ndarray x)


cdef class AbsMax:
    """Device word with the bit pattern of max |x| of the array it is attached to."""
    cdef public ndarray word
    cdef long version
    cdef long epoch
    cdef long capture
    # two persistent words {max |x|, accumulator}: the optimizer kernel rotates them itself every step
    # (sk_adam_step_split), so they also describe the array after the NEXT in-place update
    cdef public bint rotating


cdef ndarray _new_array(int ndim, const int64_t *shape, int code)
cdef SplitMat _new_split(int64_t rows, int64_t cols)
cdef void _bind_split(SplitMat m, ndarray x)
cdef AbsMax _new_rotating_absmax()
cdef long _graph_epoch()
cdef ndarray _as_device(object x)
cdef int _check(int rc) except -1
cdef float *_fptr(ndarray a) except NULL
