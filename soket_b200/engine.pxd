# engine.pxd -- Tensor type shared with nn / optim.
from soket_b200._core cimport ndarray


cdef class Tensor:
    cdef ndarray _d                # the array, or None while the tensor is DEFERRED (lazy mode)
    cdef public object _lz         # deferred elementwise node: (kind, op, inputs, scalar, rev) | None
    cdef public tuple _lshape      # its shape while deferred
    cdef public int _nuse          # deferred consumers of this deferred node
    cdef public object _dtype
    cdef public object _device
    cdef public object _grad
    cdef public bint _requires_grad
    cdef public bint _retain_grad
    cdef public object _op
    cdef public tuple _inputs
    cdef public list _partials
    cdef public int _pending
    cdef public int _visit
    cdef public int _state
    cdef public int _nedges        # partial adjoints this node receives in the running backward
    cdef public object _grad_buf   # ndarray | None: where a fused backward writes this leaf's gradient
