"""Stand-in for the part of Theano the reference's sources use (test infrastructure; see oracle/refrun/__init__.py).

A lazy expression graph over torch CPU tensors: every `Variable` holds a python function and its inputs; `function(inputs,
outputs)` binds the placeholders and evaluates the outputs with memoisation, so one call evaluates every node once, the way
a compiled Theano function would.  float32 throughout (`config.floatX`), `grad` is torch autograd.
"""
import builtins

import numpy as np
import torch


class _Config(object):
    floatX = 'float32'


config = _Config()

_TORCH = {'float32': torch.float32, 'float64': torch.float64, 'int8': torch.int8, 'int32': torch.int32, 'int64': torch.int64,
          'uint8': torch.uint8, 'bool': torch.bool}


def _dtype(d):
    return _TORCH[str(np.dtype(d))] if not isinstance(d, torch.dtype) else d


def evaluate(x, env):
    """Value of an expression (a Variable, or a python container / slice holding Variables) under `env`."""
    if isinstance(x, Variable):
        k = id(x)
        if k not in env:
            env[k] = x._compute(env)
        return env[k]
    if isinstance(x, (list, tuple)):
        return type(x)(evaluate(e, env) for e in x)
    if isinstance(x, slice):
        return slice(evaluate(x.start, env), evaluate(x.stop, env), evaluate(x.step, env))
    if isinstance(x, dict):
        return dict((k, evaluate(v, env)) for k, v in x.items())
    return x


def _t(v):
    """torch view of a python / numpy constant."""
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(v))
    return v


class Variable(object):
    _is_nonzero = True          # Theano: every variable is truthy except comparison results

    def __init__(self, fn, inputs=(), name=None, ndim=None, base=None, index=None):
        self.fn, self.inputs, self.name, self._ndim = fn, tuple(inputs), name, ndim
        self.base, self.index = base, index       # set on x[idx] results, read by set_subtensor

    def _compute(self, env):
        return self.fn(*[evaluate(i, env) for i in self.inputs])

    # ---- static information ----
    @property
    def ndim(self):
        if self._ndim is None:
            raise TypeError('ndim of %r is not known statically in this stand-in' % self.name)
        return self._ndim

    @property
    def shape(self):
        sh = Variable(lambda v: tuple(int(s) for s in v.shape), [self], name='shape', ndim=1)
        sh._len = self._ndim          # `b, c, rows, cols = x.shape` (models/FCDenseNet.py:135) unpacks it
        return sh

    def __bool__(self):
        if self._is_nonzero:
            return True
        raise TypeError('Variables do not support boolean operations.')

    def __hash__(self):
        return id(self)

    # ---- arithmetic ----
    def _bin(self, other, f, swap=False):
        nd = self._ndim
        if isinstance(other, Variable) and other._ndim is not None and nd is not None:
            nd = max(nd, other._ndim)
        elif isinstance(other, np.ndarray) and nd is not None:
            nd = max(nd, other.ndim)
        if swap:
            return Variable(lambda a, b: f(_t(b), a), [self, other], ndim=nd)
        return Variable(lambda a, b: f(a, _t(b)), [self, other], ndim=nd)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, _div)
    def __rtruediv__(self, o): return self._bin(o, _div, True)
    def __floordiv__(self, o): return self._bin(o, lambda a, b: a // b)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._bin(o, lambda a, b: a ** b, True)
    def __neg__(self): return Variable(lambda a: -a, [self], ndim=self._ndim)

    def _cmp(self, o, f):
        v = self._bin(o, lambda a, b: f(a, b).to(torch.int8))
        v._is_nonzero = False
        return v

    def __lt__(self, o): return self._cmp(o, lambda a, b: a < b)
    def __le__(self, o): return self._cmp(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._cmp(o, lambda a, b: a > b)
    def __ge__(self, o): return self._cmp(o, lambda a, b: a >= b)

    # ---- indexing ----
    def __getitem__(self, idx):
        nd = None
        if self._ndim is not None:
            items = idx if isinstance(idx, tuple) else (idx,)
            if all(isinstance(i, (slice, int, type(None), np.integer)) for i in items):
                nd = self._ndim + sum(i is None for i in items) - sum(isinstance(i, (int, np.integer)) for i in items)
            elif len(items) == self._ndim and all(isinstance(i, Variable) for i in items):
                nd = 1
        return Variable(_getitem, [self, idx], ndim=nd, base=self, index=idx)

    def __iter__(self):
        n = getattr(self, '_len', None)
        if n is None:
            raise TypeError('iteration over a symbolic variable of unknown length')
        return iter([self[i] for i in range(n)])

    # ---- methods the reference calls ----
    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])
        return Variable(lambda a: _dimshuffle(a, pattern), [self], ndim=len(pattern))

    def reshape(self, shape, ndim=None):
        return Variable(lambda a, s: a.reshape(tuple(int(e) for e in s)), [self, tuple(shape)], ndim=len(shape))

    def flatten(self, ndim=1):
        assert ndim == 1
        return Variable(lambda a: a.reshape(-1), [self], ndim=1)

    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (list, tuple)):
            axes = tuple(axes[0])
        return Variable(lambda a: a.permute(*axes), [self], ndim=len(axes))

    def astype(self, dtype):
        return Variable(lambda a: a.to(_dtype(dtype)), [self], ndim=self._ndim)

    def _reduce(self, f, axis, keepdims):
        nd = self._ndim
        if nd is not None:
            nd = 0 if axis is None else nd - (0 if keepdims else (len(axis) if isinstance(axis, (list, tuple)) else 1))
        return Variable(lambda a: f(a, axis, keepdims), [self], ndim=nd)

    def sum(self, axis=None, keepdims=False):
        return self._reduce(lambda a, ax, k: (a.sum() if ax is None else a.sum(dim=ax, keepdim=k)), axis, keepdims)

    def mean(self, axis=None, keepdims=False):
        return self._reduce(lambda a, ax, k: (a.mean() if ax is None else a.mean(dim=ax, keepdim=k)), axis, keepdims)

    def var(self, axis=None, keepdims=False):          # Theano's var is the biased (population) variance
        return self._reduce(lambda a, ax, k: (a.var(unbiased=False) if ax is None else a.var(dim=ax, unbiased=False, keepdim=k)),
                            axis, keepdims)

    def max(self, axis=None, keepdims=False):
        return self._reduce(lambda a, ax, k: (a.max() if ax is None else a.amax(dim=ax, keepdim=k)), axis, keepdims)

    def min(self, axis=None, keepdims=False):
        return self._reduce(lambda a, ax, k: (a.min() if ax is None else a.amin(dim=ax, keepdim=k)), axis, keepdims)

    def exp(self): return Variable(torch.exp, [self], ndim=self._ndim)
    def diagonal(self): return Variable(lambda a: a.diagonal(), [self], ndim=1)

    def nonzero(self):
        assert self._ndim is not None
        return tuple(Variable(lambda a, d=d: a.nonzero(as_tuple=True)[d], [self], ndim=1) for d in range(self._ndim))

    def eval(self, inputs_to_values=None):
        env = {}
        for k, v in (inputs_to_values or {}).items():
            env[id(k)] = torch.as_tensor(np.asarray(v))
        out = evaluate(self, env)
        return out.detach().numpy() if isinstance(out, torch.Tensor) else np.asarray(out)


def _div(a, b):
    """Theano `/` on tensors is true division; integer / integer gives floatX-or-wider, never floor."""
    ai = isinstance(a, torch.Tensor) and not a.dtype.is_floating_point
    bi = (isinstance(b, torch.Tensor) and not b.dtype.is_floating_point) or isinstance(b, int)
    if ai and bi:
        return a.to(torch.float64) / b
    return a / b


def _getitem(v, idx):
    if isinstance(v, tuple):            # a shape
        return v[idx]
    return v[idx]


def _dimshuffle(a, pattern):
    keep = [p for p in pattern if p != 'x']
    dropped = [d for d in range(a.dim()) if d not in keep]
    for d in dropped:
        assert a.shape[d] == 1, 'dimshuffle drops a non-broadcastable axis'
    a = a.permute(*(keep + dropped)).reshape([a.shape[d] for d in keep])
    for i, p in enumerate(pattern):
        if p == 'x':
            a = a.unsqueeze(i)
    return a


class Placeholder(Variable):
    def __init__(self, ndim, dtype, name=None):
        Variable.__init__(self, None, (), name=name, ndim=ndim)
        self.dtype = dtype

    def _compute(self, env):
        raise ValueError('input %r of the graph was not given a value' % (self.name,))


class SharedVariable(Variable):
    def __init__(self, value, name=None):
        value = np.array(value)
        Variable.__init__(self, None, (), name=name, ndim=value.ndim)
        self.value = value

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, value, borrow=False):
        value = np.asarray(value)
        self.value = np.array(value, dtype=self.value.dtype if self.value.dtype.kind == 'f' else value.dtype)

    @property
    def dtype(self):
        return str(self.value.dtype)

    def _compute(self, env):
        t = torch.tensor(self.value)
        if t.dtype.is_floating_point:
            t.requires_grad_(True)
        return t


def shared(value, name=None, **kwargs):
    return SharedVariable(value, name=name)


class Function(object):
    """`theano.function(inputs, outputs, updates=...)`: numpy in, numpy out."""

    def __init__(self, inputs, outputs, updates=None, **kwargs):
        self.inputs, self.outputs = list(inputs), outputs
        self.updates = list(updates.items()) if hasattr(updates, 'items') else list(updates or [])

    def __call__(self, *args):
        assert len(args) == len(self.inputs), 'expected %d inputs, got %d' % (len(self.inputs), len(args))
        from .sandbox import rng_mrg
        rng_mrg.STATE['call'] += 1          # random draws are logged with the function call that consumed them
        env = {}
        for var, val in zip(self.inputs, args):
            t = torch.tensor(np.asarray(val), dtype=_dtype(var.dtype))
            if t.dtype.is_floating_point:
                t.requires_grad_(True)          # so that T.grad can differentiate sub-expressions of this call
            env[id(var)] = t
        outs = evaluate(self.outputs, env)
        new = [(sv, evaluate(expr, env)) for sv, expr in self.updates]
        for sv, val in new:
            sv.set_value(val.detach().numpy())

        def conv(o):
            if isinstance(o, torch.Tensor):
                return o.detach().numpy().copy()
            if isinstance(o, (list, tuple)):
                return [conv(e) for e in o]
            return np.asarray(o)
        return conv(outs)


def function(inputs, outputs=None, updates=None, **kwargs):
    return Function(inputs, outputs, updates, **kwargs)


def grad(cost, wrt, known_grads=None, **kwargs):
    """theano.grad.  `known_grads` = {expression: its gradient} seeds the backward pass (layers/mylayers.py:111-112,
    lasagne InverseLayer)."""
    single = not isinstance(wrt, (list, tuple))
    wrts = [wrt] if single else list(wrt)
    srcs, seeds = [], []
    if cost is not None:
        srcs.append(cost)
        seeds.append(None)
    for k, v in (known_grads or {}).items():
        srcs.append(k)
        seeds.append(v)

    def run(ws, ss, gs):
        gs = [torch.ones_like(s) if g is None else _t(g).to(s.dtype).expand_as(s) for s, g in zip(ss, gs)]
        out = torch.autograd.grad(ss, ws, grad_outputs=gs, retain_graph=True, allow_unused=True)
        return [torch.zeros_like(w) if o is None else o for w, o in zip(ws, out)]
    allg = Variable(run, [wrts, srcs, seeds])
    res = [Variable(lambda g, i=i: g[i], [allg], ndim=w._ndim) for i, w in enumerate(wrts)]
    return res[0] if single else res


from . import gradient, tensor  # noqa: E402,F401
