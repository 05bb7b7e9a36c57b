"""theano.tensor.extra_ops stand-in (test infrastructure)."""
import torch

from theano import Variable


def repeat(x, repeats, axis=None):
    return Variable(lambda a: torch.repeat_interleave(a, int(repeats), dim=axis), [x], ndim=x._ndim)
