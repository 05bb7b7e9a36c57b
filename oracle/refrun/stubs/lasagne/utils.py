import numpy as np
import theano


def floatX(arr):
    return np.asarray(arr, dtype=theano.config.floatX)


def as_tuple(x, N, t=None):
    try:
        X = tuple(x)
    except TypeError:
        X = (x,) * N
    if t is not None and not all(isinstance(v, t) for v in X):
        raise TypeError('expected a single value or an iterable of %s' % t.__name__)
    if len(X) != N:
        raise ValueError('expected a single value or an iterable with length %d' % N)
    return X


def unique(l):
    new, seen = [], set()
    for el in l:
        if id(el) not in seen:
            new.append(el)
            seen.add(id(el))
    return new
