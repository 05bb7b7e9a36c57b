"""theano.tensor stand-in: the functions the reference's sources call (test infrastructure)."""
import numpy as np
import torch

from .. import Placeholder, Variable, _dtype, _t, config, grad  # noqa: F401
from . import extra_ops, nnet  # noqa: F401

TensorVariable = Variable


def _placeholder(ndim):
    def make(name=None, dtype=None):
        return Placeholder(ndim, dtype or config.floatX, name)
    return make


scalar, vector, matrix, tensor3, tensor4 = [_placeholder(n) for n in range(5)]
ivector = lambda name=None: Placeholder(1, 'int32', name)          # noqa: E731
imatrix = lambda name=None: Placeholder(2, 'int32', name)          # noqa: E731
itensor3 = lambda name=None: Placeholder(3, 'int32', name)         # noqa: E731


def _nd(x):
    if isinstance(x, Variable):
        return x._ndim
    return np.ndim(x)


def as_tensor_variable(x):
    if isinstance(x, Variable):
        return x
    return Variable(lambda: _t(np.asarray(x)), [], ndim=np.ndim(x))


def _tensor(v, like=None):
    """torch tensor from whatever an evaluated argument is (tensor, shape tuple, python number)."""
    if isinstance(v, torch.Tensor):
        return v
    return torch.as_tensor(np.asarray(v))


def _reduce(name):
    def f(x, axis=None, keepdims=False, **kw):
        if isinstance(x, Variable) and x._ndim == 1 and x.name == 'shape' or _is_shape(x):
            fn = {'sum': np.sum, 'prod': np.prod, 'max': np.max, 'min': np.min}[name]
            return Variable(lambda v: int(fn(np.asarray(v))), [x], ndim=0)
        if name == 'prod':
            return Variable(lambda v: _tensor(v).prod() if axis is None else _tensor(v).prod(dim=axis, keepdim=keepdims), [x],
                            ndim=0 if axis is None else None)
        return getattr(as_tensor_variable(x), name)(axis=axis, keepdims=keepdims)
    return f


def _is_shape(x):
    """x[a:b] of a shape variable."""
    return isinstance(x, Variable) and x.base is not None and x.base.name == 'shape'


sum = _reduce('sum')
prod = _reduce('prod')
max = _reduce('max')
min = _reduce('min')
mean = lambda x, axis=None, keepdims=False: x.mean(axis=axis, keepdims=keepdims)      # noqa: E731


def shape(x):
    return x.shape


def _cmp(f):
    def g(a, b):
        a = as_tensor_variable(a)
        v = a._bin(b, lambda x, y: f(x, y).to(torch.int8))
        return v
    return g


eq = _cmp(lambda a, b: a == b)
neq = _cmp(lambda a, b: a != b)
lt = _cmp(lambda a, b: a < b)
gt = _cmp(lambda a, b: a > b)


def argmax(x, axis=None, keepdims=False):
    return Variable(lambda a: a.argmax(dim=axis, keepdim=keepdims), [x], ndim=None if x._ndim is None else x._ndim - (0 if keepdims else 1))


def cast(x, dtype):
    return as_tensor_variable(x).astype(dtype)


def clip(x, lo, hi):
    return Variable(lambda a, l, h: torch.clamp(a, l, h), [x, lo, hi], ndim=_nd(x))


def log(x): return Variable(torch.log, [x], ndim=_nd(x))
def exp(x): return Variable(torch.exp, [x], ndim=_nd(x))
def sqrt(x): return Variable(torch.sqrt, [x], ndim=_nd(x))
def sqr(x): return x * x
def inv(x): return Variable(lambda a: 1.0 / a, [x], ndim=_nd(x))
def abs_(x): return Variable(torch.abs, [x], ndim=_nd(x))
def maximum(a, b): return as_tensor_variable(a)._bin(b, lambda x, y: torch.maximum(x, torch.as_tensor(y, dtype=x.dtype)))
def minimum(a, b): return as_tensor_variable(a)._bin(b, lambda x, y: torch.minimum(x, torch.as_tensor(y, dtype=x.dtype)))
def add(a, b): return a + b
def mul(a, b): return a * b


def zeros(shape, dtype=None):
    return Variable(lambda s: torch.zeros(tuple(int(e) for e in (s if isinstance(s, (tuple, list)) else [s])),
                                          dtype=_dtype(dtype or config.floatX)), [shape],
                    ndim=len(shape) if isinstance(shape, (tuple, list)) else None)


def ones(shape, dtype=None):
    return Variable(lambda s: torch.ones(tuple(int(e) for e in (s if isinstance(s, (tuple, list)) else [s])),
                                         dtype=_dtype(dtype or config.floatX)), [shape],
                    ndim=len(shape) if isinstance(shape, (tuple, list)) else None)


def ones_like(x, dtype=None):
    return Variable(lambda a: torch.ones_like(a, dtype=_dtype(dtype) if dtype else a.dtype), [x], ndim=_nd(x))


def zeros_like(x, dtype=None):
    return Variable(lambda a: torch.zeros_like(a, dtype=_dtype(dtype) if dtype else a.dtype), [x], ndim=_nd(x))


def flatten(x, outdim=1):
    if outdim == 1:
        return x.flatten()
    return Variable(lambda a: a.reshape(tuple(a.shape[:outdim - 1]) + (-1,)), [x], ndim=outdim)


def stack(tensors, axis=0):
    return Variable(lambda ts: torch.stack([_tensor(t) for t in ts], dim=axis), [list(tensors)], ndim=(_nd(tensors[0]) or 0) + 1
                    if _nd(tensors[0]) is not None else None)


def concatenate(tensors, axis=0):
    return Variable(lambda ts: torch.cat(list(ts), dim=axis), [list(tensors)], ndim=_nd(tensors[0]))


def set_subtensor(sub, value, **kwargs):
    """x[idx] <- value as a new tensor.  `sub` must be the result of indexing; the value is cast to x's dtype, as the
    numpy assignment inside Theano's IncSubtensor does."""
    assert isinstance(sub, Variable) and sub.base is not None, 'set_subtensor needs x[idx]'

    def run(base, idx, val):
        out = base.clone()
        out[idx] = val if not isinstance(val, torch.Tensor) else val.to(out.dtype)
        return out
    return Variable(run, [sub.base, sub.index, value], ndim=sub.base._ndim)


def inc_subtensor(sub, value, **kwargs):
    def run(base, idx, val):
        out = base.clone()
        out[idx] = out[idx] + val
        return out
    return Variable(run, [sub.base, sub.index, value], ndim=sub.base._ndim)


def dot(a, b):
    return Variable(lambda x, y: x @ y, [a, b])


def tensordot(a, b, axes):
    return Variable(lambda x, y: torch.tensordot(x, y, axes), [a, b])
