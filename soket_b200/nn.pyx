# cython: language_level=3, boundscheck=True, wraparound=True, cdivision=True
"""soket_b200.nn -- Soket's nn API (soket/nn/module.pyx, prototypes.pyx,
functional.pyx, init.py) on the fused sm_100a kernels.

Same module names, constructor arguments, parameter discovery rules and quirks
as the reference; what changes is what one forward call launches:

  Linear              matmul + add                 -> 1 GEMM with bias epilogue
  Linear -> ReLU      matmul + add + maximum       -> 1 GEMM with bias+ReLU epilogue
  LayerNorm           9 array calls                -> 1 kernel (x read once)
  LayerNorm -> ReLU   10                           -> same kernel, ReLU epilogue
  LN -> ReLU -> Dropout  10 + binomial+astype+2 mul -> same kernel, dropout in the epilogue (backward: in the prologue)
  Residual -> ReLU    add + maximum (+ LN's 9)     -> relu(x + LN(.)) in LN's epilogue
  BatchNorm1d         17 array calls               -> 3 kernels (stats, finalise, apply)
  Dropout             binomial+astype+2 mul        -> 1 kernel
  SoftmaxCrossEntropy ~12 fwd + ~8 bwd, one-hot    -> 1 kernel, labels read directly

The pairings are found by ``Sequential`` looking one module ahead.  Fusion can be
switched off (``set_fusion(False)``) to run the reference's op-by-op sequence on
the same kernels (used by the parity tests).
"""
from collections import OrderedDict
import math

from soket_b200.engine cimport Tensor
from soket_b200 import engine as E
from soket_b200 import _core as B

cdef bint _FUSE = True


def set_fusion(bint on):
    global _FUSE
    _FUSE = on


def fusion_enabled():
    return _FUSE


# ============================================================================ Module
def _walk_modules(value):
    """soket/nn/module.pyx:12-41."""
    if isinstance(value, Module):
        yield value
        for s in (<Module> value)._storage:
            yield from _walk_modules(s)
        yield from _walk_modules(value.__dict__)
    elif type(value) is dict:
        for v in value.values():
            yield from _walk_modules(v)
    elif type(value) in (list, tuple):
        for v in value:
            yield from _walk_modules(v)


def _walk_params(value):
    """soket/nn/module.pyx:44-73."""
    if isinstance(value, Tensor):
        yield value
    elif isinstance(value, Module):
        for s in (<Module> value)._storage:
            yield from _walk_params(s)
        yield from _walk_params(value.__dict__)
    elif type(value) is dict:
        for v in value.values():
            yield from _walk_params(v)
    elif type(value) in (list, tuple):
        for v in value:
            yield from _walk_params(v)


cdef void _set_train(Module m, bint mode):
    """soket/nn/module.pyx:76-89."""
    m._training = mode
    for s in m._storage:
        if isinstance(s, Module):
            _set_train(<Module> s, mode)
    for v in (<object> m).__dict__.values():
        if isinstance(v, Module):
            _set_train(<Module> v, mode)


cdef class Module:
    """soket/nn/module.pyx:94-217.  `_storage` holds a builtin module's children /
    parameters; Python-subclass attributes live in `__dict__`; both are walked by
    parameters() / modules() / train()."""
    cdef public bint _training
    cdef public list _storage
    cdef public bint _builtin
    cdef dict __dict__

    def __init__(self):
        self._training = True
        self._storage = []
        self._builtin = False

    @property
    def training(self):
        return self._training

    def modules(self):
        return _walk_modules(self)

    def parameters(self):
        return _walk_params(self)

    def train(self, mode=True):
        _set_train(self, bool(mode))

    def eval(self):
        _set_train(self, False)

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return None

    def __call__(self, *args):
        if len(args) == 0:
            raise ValueError('Expected atleast one positional argument!')
        if self._builtin:
            return self._fast_forward(args[0], None if len(args) == 1 else args[1])
        return self.forward(*args)

    def __str__(self):
        if type(self) is Module:
            return 'soket.nn.Module()'
        return self.__class__.__name__ + '()'

    def __repr__(self):
        return self.__str__()


cdef class Identity(Module):
    def forward(self, X):
        return X

    def __str__(self):
        return 'soket.nn.Identity()'


# ============================================================================ layers
cdef class Linear(Module):
    """Y = X @ W + b, W stored (in, out)  (soket/nn/prototypes.pyx:36-138)."""
    cdef public object _feature_in, _feature_out

    def __init__(self, feature_in, feature_out, bias=True, device=None, dtype=None):
        Module.__init__(self)
        self._builtin = True
        self._feature_in = feature_in
        self._feature_out = feature_out
        self._storage = [
            E.zeros((feature_in, feature_out), dtype=dtype, requires_grad=True),
            E.zeros((feature_out,), dtype=dtype, requires_grad=True) if bias else None,
        ]

    @property
    def weight(self):
        return self._storage[0]

    @property
    def bias(self):
        return self._storage[1]

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return self._forward(X, False)

    cpdef Tensor _forward(self, Tensor X, bint relu):
        cdef Tensor W = self._storage[0]
        if _FUSE and X._dtype.name == 'float32' and W._dtype.name == 'float32':
            return E.linear(X, W, self._storage[1], relu)
        Y = X @ W
        if self._storage[1] is not None:
            Y = Y + self._storage[1]
        return E.relu_(Y) if relu else Y

    def __str__(self):
        return (f'soket.nn.Linear(feature_in={self._feature_in}, feature_out={self._feature_out}, '
                f'bias={self.bias is not None}, dtype={self.weight.dtype.name}, device=GPU:0)')


cdef class ReLU(Module):
    def __init__(self):
        Module.__init__(self)
        self._builtin = True

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return E.relu_(X)

    def __str__(self):
        return 'soket.nn.ReLU()'


cdef class LayerNorm(Module):
    """soket/nn/prototypes.pyx:633-722."""
    cdef public tuple _normalized_shape
    cdef public object _eps

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True, bias=True,
                 device=None, dtype=None):
        Module.__init__(self)
        self._builtin = True
        shp = (normalized_shape,) if type(normalized_shape) is int else tuple(normalized_shape)
        self._normalized_shape = shp
        self._eps = eps
        self._storage = [None, None]
        if elementwise_affine is True:
            self._storage[0] = E.ones(shp, dtype=dtype, requires_grad=True)
            self._storage[1] = None if bias is False else E.zeros(shp, dtype=dtype, requires_grad=True)

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return self._forward(X, False, None)

    cpdef Tensor _forward(self, Tensor X, bint relu, object residual):
        if _FUSE:
            return E.layer_norm(X, self._storage[0], self._storage[1], self._eps, relu, residual)
        out = E._layer_norm_unfused(X, self._storage[0], self._storage[1], self._eps, False, None)
        if residual is not None:
            out = residual + out
        return E.relu_(out) if relu else out

    def __str__(self):    # prototypes.pyx:702-720
        gamma = self._storage[0]
        affine = gamma is not None and gamma.requires_grad
        dtype = gamma.dtype.name if gamma is not None else 'float32'
        return (f'soket.nn.LayerNorm({self._normalized_shape}, eps={self._eps}, elementwise_affine={affine}, '
                f'dtype={dtype}, device=GPU:0)')


cdef class _BatchNormBase(Module):
    """soket/nn/prototypes.pyx:493-598."""
    cdef public object _eps, _momentum
    cdef public object _running_mean, _running_var
    cdef public object _num_features

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 device=None, dtype=None):
        Module.__init__(self)
        self._builtin = True
        self._eps = eps
        self._momentum = momentum
        self._num_features = num_features
        self._storage = [
            E.ones((num_features,), dtype=dtype, requires_grad=affine),
            E.zeros((num_features,), dtype=dtype, requires_grad=affine),
        ]
        if track_running_stats is True:
            # 0-d tensors that become (1, C) after the first step (prototypes.pyx:550-551)
            self._running_mean = Tensor(0.0, None, E._dt(dtype))
            self._running_var = Tensor(1.0, None, E._dt(dtype))
        else:
            self._running_mean = None
            self._running_var = None

    @property
    def running_mean(self):
        return self._running_mean

    @property
    def running_var(self):
        return self._running_var

    def _check_input_dim(self, X):
        raise NotImplementedError

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return self._forward(X, False)

    cpdef Tensor _forward(self, Tensor X, bint relu):
        self._check_input_dim(X)
        if _FUSE:
            return E.batch_norm(X, self._running_mean, self._running_var, self._storage[0],
                                self._storage[1], self._training, self._momentum, self._eps, relu)
        nd = X.ndim
        axes = (0,) + tuple(range(2, nd))
        obs = X.shape[0]
        for s in X.shape[2:]:
            obs *= s
        out = E._norm_unfused(X, self._storage[0], self._storage[1], axes, obs, self._eps, False,
                              self._running_mean, self._running_var, self._momentum)
        return E.relu_(out) if relu else out

    def __str__(self):
        return (f'soket.nn.{self.__class__.__name__}({self._num_features}, eps={self._eps}, '
                f'momentum={self._momentum})')


cdef class BatchNorm1d(_BatchNormBase):
    def _check_input_dim(self, X):
        if X.ndim != 2 and X.ndim != 3:
            raise ValueError('Expected 2D or 3D input tensor!')


cdef class BatchNorm2d(_BatchNormBase):
    def _check_input_dim(self, X):
        if X.ndim != 4:
            raise ValueError('Expected a 4D tensor!')


cdef class BatchNorm3d(_BatchNormBase):
    def _check_input_dim(self, X):
        if X.ndim != 5:
            raise ValueError('Expected a 5D input tensor!')


cdef class Dropout(Module):
    """soket/nn/prototypes.pyx:729-771."""
    cdef public object _keep_rate, _r_keep_rate

    def __init__(self, p=0.5):
        Module.__init__(self)
        self._builtin = True
        self._keep_rate = 1.0 - p
        self._r_keep_rate = 1.0 / (1.0 - p)

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        if not self._training:
            return X
        if self._keep_rate == 1.0:
            # binomial(1, 1.0) is all ones and x * 1 * 1.0 == x exactly: skip the pass
            return X
        if _FUSE:
            return E.dropout(X, self._keep_rate)
        mask = E.randb(X.shape, p=self._keep_rate, dtype=X.dtype)
        return (X * mask) * self._r_keep_rate

    def __str__(self):
        return f'soket.nn.Dropout(p={1.0 - self._keep_rate})'


cdef class Residual(Module):
    """soket/nn/prototypes.pyx:256-284.  Quirk Q1: the wrapped layers are a private
    attribute -- NOT in `_storage` nor `__dict__` -- so parameters() / modules() /
    train() never reach inside (inner Linear weights are invisible to the
    optimiser and to kaiming init unless the user keeps another reference)."""
    cdef Module _layers

    def __init__(self, layers):
        Module.__init__(self)
        self._builtin = True
        self._layers = layers

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return X + self._layers(X)

    cpdef Tensor _forward_relu(self, Tensor X):
        """relu(X + layers(X)) with the add+ReLU pushed into the last layer when it is
        a LayerNorm, else one add+relu kernel."""
        cdef Module inner = self._layers
        cdef Sequential seq
        if _FUSE and isinstance(inner, Sequential) and inner._builtin:
            seq = <Sequential> inner
            if len(seq._storage) and type(seq._storage[-1]) is LayerNorm:
                h = seq._run(X, len(seq._storage) - 1)
                return (<LayerNorm> seq._storage[-1])._forward(h, True, X)
        if _FUSE:
            return E.add_relu(X, inner(X))
        return E.relu_(X + inner(X))

    def __str__(self):
        return f'soket.nn.Residual(layers={self._layers})'


cdef class Sequential(Module):
    """soket/nn/prototypes.pyx:143-253, plus one-module look-ahead fusion."""
    cdef public list _odict_keys

    def __init__(self, *modules):
        Module.__init__(self)
        self._builtin = True
        self._odict_keys = []
        if len(modules) > 0:
            first = modules[0]
            if isinstance(first, OrderedDict):
                self._odict_keys = list(first.keys())
                modules = tuple(first.values())
            elif type(first) is tuple:
                modules = first
        for m in modules:
            if not isinstance(m, Module):
                raise ValueError(f'{type(m)} is not a Module subclass!')
        self._storage = list(modules)

    def append(self, Module module):
        self._storage.append(module)
        return self

    cpdef Tensor _run(self, Tensor X, int stop):
        """Run modules [0, stop) with look-ahead fusion."""
        cdef Tensor Y = X
        cdef int i = 0
        cdef Module m
        cdef object nxt
        while i < stop:
            m = <Module> self._storage[i]
            nxt = self._storage[i + 1] if i + 1 < stop else None
            if _FUSE and type(nxt) is ReLU and m._builtin:
                if type(m) is LayerNorm and i + 2 < stop and type(self._storage[i + 2]) is Dropout:
                    # LayerNorm - ReLU - Dropout (model.py:24-31): one kernel
                    drop = <Dropout> self._storage[i + 2]
                    if drop._training and drop._keep_rate < 1.0:
                        ln = <LayerNorm> m
                        Y = E.layer_norm_dropout(Y, ln._storage[0], ln._storage[1], ln._eps, True, drop._keep_rate)
                        i += 3; continue
                if type(m) is Linear:
                    Y = (<Linear> m)._forward(Y, True); i += 2; continue
                if type(m) is LayerNorm:
                    Y = (<LayerNorm> m)._forward(Y, True, None); i += 2; continue
                if isinstance(m, _BatchNormBase):
                    Y = (<_BatchNormBase> m)._forward(Y, True); i += 2; continue
                if type(m) is Residual:
                    Y = (<Residual> m)._forward_relu(Y); i += 2; continue
            if m._builtin:
                Y = m._fast_forward(Y, None)
            else:
                Y = m.forward(Y)
            i += 1
        return Y

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        return self._run(X, len(self._storage))

    def __str__(self):
        res = 'soket.nn.Sequential('
        for i, m in enumerate(self._storage):
            if i == 0:
                res += '\n'
            lines = str(m).splitlines()
            key = self._odict_keys[i] if i < len(self._odict_keys) else str(i)
            res += '  (' + key + '): ' + '\n'.join(lines[:1] + ['  ' + x for x in lines[1:]]) + '\n'
        return res + ')'


cdef class SoftmaxCrossEntropyLoss(Module):
    """soket/nn/prototypes.pyx:329-486."""
    cdef public str _reduction

    def __init__(self, reduction='mean'):
        Module.__init__(self)
        self._builtin = True
        if reduction not in ('sum', 'mean', 'none'):
            raise ValueError(f'Invalid reduction type - {reduction}')
        self._reduction = reduction

    cpdef Tensor _fast_forward(self, Tensor X, object y):
        if X is None or y is None:
            raise ValueError('Expected tensors as inputs, got None instead')
        cdef Tensor t = <Tensor> y
        if X.ndim == 0:
            raise ValueError('Expected the classes tensor to be atleast 1D!')
        if X.ndim == 1:
            if t.ndim >= 1:
                raise ValueError('Incompatible targets tensor, expected shape to be ()')
        elif t.ndim + 1 != X.ndim or X.shape[0] != t.shape[0] or X.shape[2:] != t.shape[1:]:
            raise ValueError(f'Incompatible targets tensor shape - {X.shape} and {t.shape}')
        if (_FUSE and self._reduction == 'mean' and X.ndim == 2 and X._dtype.name == 'float32'
                and t._dtype.name not in ('float16', 'float32', 'float64')):
            return E.softmax_cross_entropy(X, t)
        return self._unfused(X, t)

    def _unfused(self, Tensor X, Tensor t):
        """The reference's op sequence (prototypes.pyx:393-474, forward.pyx:250-271)."""
        r_axis = 1 if X.ndim >= 2 else 0
        onehot = E.one_hot(t, X.shape[r_axis], dtype=X.dtype)._data
        if t.ndim > 2:
            perm = (0, X.ndim - 1) + tuple(range(1, X.ndim - 1))
            onehot = B.transpose(onehot, perm)
        return E.softmax_cross_entropy_unfused(X, onehot, (r_axis,), self._reduction)

    def __str__(self):
        return f"soket.nn.SoftmaxCrossEntropyLoss(reduction='{self._reduction}')"


# ============================================================================ init (soket/nn/init.py)
class init:
    """soket/nn/init.py.  Quirk Q9: the reference passes the VARIANCE where randn
    expects the std, and a squared gain as the uniform bound; kept."""

    # soket/nn/init.py:41-48 (measured gains, e.g. relu 1.4142 -- not sqrt(2))
    _GAINS = {'linear': 1.0, 'identity': 1.0, 'conv': 1.0, 'sigmoid': 1.0, 'tanh': 1.6666, 'relu': 1.4142}

    @staticmethod
    def _prologue(shape, mode, nonlinearity):
        assert mode == 'fan_in' or mode == 'fan_out', 'Invalid mode'
        assert init._GAINS.get(nonlinearity) is not None, 'Invalid nonlinearity'
        fan = {'fan_out': shape[-1], 'fan_in': shape[-2]}
        return init._GAINS[nonlinearity], fan[mode]

    @staticmethod
    def xavier_normal(tensor, gain=1.0):
        fan_in, fan_out = tensor.shape[-2:]
        std_sq = gain * gain * (2 / (fan_in + fan_out))
        tensor.data = E.randn(tensor.shape, mean=0.0, std=std_sq, dtype=tensor.dtype)

    @staticmethod
    def xavier_uniform(tensor, gain=1.0):
        fan_in, fan_out = tensor.shape[-2:]
        a = gain * gain * (6 / (fan_in + fan_out))
        tensor.data = E.rand(tensor.shape, low=-a, high=a, dtype=tensor.dtype)

    @staticmethod
    def kaiming_normal(tensor, mode='fan_in', nonlinearity='relu'):
        g, fan = init._prologue(tensor.shape, mode, nonlinearity)
        tensor.data = E.randn(tensor.shape, mean=0.0, std=g * g / fan, dtype=tensor.dtype)

    @staticmethod
    def kaiming_uniform(tensor, mode='fan_in', nonlinearity='relu'):
        g, fan = init._prologue(tensor.shape, mode, nonlinearity)
        bound = g * g * 3 / fan
        tensor.data = E.rand(tensor.shape, low=-bound, high=bound, dtype=tensor.dtype)



# soket.nn.init and soket.nn.functional are modules in the reference (`from soket.nn.init import
# kaiming_normal`, `soket.nn.functional.layer_norm`); this file is one extension module, so the two
# namespaces are registered as importable modules of it.
def _submodule(name, members):
    import sys
    import types
    mod = types.ModuleType('soket_b200.nn.' + name)
    mod.__dict__.update(members)
    sys.modules['soket_b200.nn.' + name] = mod
    return mod


functional = _submodule('functional', {'batch_norm': E.batch_norm, 'layer_norm': E.layer_norm})
_submodule('init', {k: getattr(init, k) for k in ('xavier_normal', 'xavier_uniform', 'kaiming_normal',
                                                   'kaiming_uniform')})

kaiming_normal = init.kaiming_normal
kaiming_uniform = init.kaiming_uniform
xavier_normal = init.xavier_normal
xavier_uniform = init.xavier_uniform
