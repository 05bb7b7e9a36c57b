"""MRG_RandomStreams stand-in (test infrastructure).

Theano's MRG31k3p stream itself is not reproduced -- what is reproduced is its STRUCTURE: every symbolic `normal()` call is
an independent stream (Theano allocates a fresh rstate per call and never merges random nodes), evaluated once per function
call.  Values come from torch: the k-th draw evaluated in the process is randn(shape, generator seeded with SEED0 + k), so a
test can regenerate exactly the numbers the reference consumed from the log below.
"""
import torch

from theano import Variable

SEED0 = 987650000
STATE = {'created': 0, 'evaluated': 0, 'call': 0, 'log': []}


def draw(k, shape):
    """The k-th evaluated draw: standard normal, float32, `shape`."""
    return torch.randn(tuple(int(s) for s in shape), generator=torch.Generator().manual_seed(SEED0 + int(k)), dtype=torch.float32)


class MRG_RandomStreams(object):
    def __init__(self, seed=None, **kwargs):
        self.seed = seed

    def normal(self, size, avg=0.0, std=1.0, **kwargs):
        created = STATE['created']
        STATE['created'] += 1

        def run(shape):
            k = STATE['evaluated']
            STATE['evaluated'] += 1
            STATE['log'].append({'k': k, 'created': created, 'call': STATE['call'], 'shape': tuple(int(s) for s in shape)})
            return draw(k, shape) * float(std) + float(avg)
        return Variable(run, [size], name='normal')

    def uniform(self, *args, **kwargs):
        raise NotImplementedError('only normal() streams are used by the reference path')

    binomial = uniform
