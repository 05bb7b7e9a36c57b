"""lasagne.init stand-in: the initialisers are only drawn from before a checkpoint overwrites them."""
import numpy as np

from .random import get_rng
from .utils import floatX


class Initializer(object):
    def __call__(self, shape):
        return self.sample(shape)

    def sample(self, shape):
        raise NotImplementedError()


class Constant(Initializer):
    def __init__(self, val=0.0):
        self.val = val

    def sample(self, shape):
        return floatX(np.ones(shape) * self.val)


class Uniform(Initializer):
    def __init__(self, range=0.01, std=None, mean=0.0):
        self.range = (-range, range) if not isinstance(range, tuple) else range

    def sample(self, shape):
        return floatX(get_rng().uniform(low=self.range[0], high=self.range[1], size=shape))


class Glorot(Initializer):
    def __init__(self, gain=1.0, c01b=False):
        self.gain = np.sqrt(2) if gain == 'relu' else gain

    def sample(self, shape):
        n1, n2 = shape[:2]
        rf = np.prod(shape[2:])
        std = self.gain * np.sqrt(2.0 / ((n1 + n2) * rf))
        a = np.sqrt(3) * std
        return floatX(get_rng().uniform(low=-a, high=a, size=shape))


GlorotUniform = Glorot


class He(Initializer):
    def __init__(self, gain=1.0, c01b=False):
        self.gain = np.sqrt(2) if gain == 'relu' else gain

    def sample(self, shape):
        fan_in = np.prod(shape[1:]) if len(shape) > 2 else shape[0]
        a = np.sqrt(3) * self.gain * np.sqrt(1.0 / fan_in)
        return floatX(get_rng().uniform(low=-a, high=a, size=shape))


HeUniform = He
