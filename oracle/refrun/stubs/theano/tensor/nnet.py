"""theano.tensor.nnet stand-in (test infrastructure)."""
import torch

from theano import Variable


def softmax(x):
    """Row softmax of a matrix, computed the way Theano's Softmax op does: exp(x - max) / sum."""
    def run(a):
        assert a.dim() == 2
        e = torch.exp(a - a.max(dim=1, keepdim=True).values)
        return e / e.sum(dim=1, keepdim=True)
    return Variable(run, [x], ndim=2)


def relu(x, alpha=0):
    assert alpha == 0
    return Variable(lambda a: 0.5 * (a + torch.abs(a)), [x], ndim=x._ndim)


def sigmoid(x):
    return Variable(torch.sigmoid, [x], ndim=x._ndim)


def categorical_crossentropy(coding_dist, true_dist):
    def run(p, t):
        if t.dim() == p.dim():
            return -(t * torch.log(p)).sum(dim=1)
        return -torch.log(p[torch.arange(p.shape[0]), t.long()])
    return Variable(run, [coding_dist, true_dist], ndim=1)


def binary_crossentropy(output, target):
    return Variable(lambda o, t: -(t * torch.log(o) + (1.0 - t) * torch.log(1.0 - o)), [output, target], ndim=output._ndim)
