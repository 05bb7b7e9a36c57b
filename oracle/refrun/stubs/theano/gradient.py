"""theano.gradient stand-in (test infrastructure)."""
from . import Variable, grad  # noqa: F401


def zero_grad(x):
    return Variable(lambda a: a.detach(), [x], ndim=x._ndim)


disconnected_grad = zero_grad
