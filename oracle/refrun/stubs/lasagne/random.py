import numpy as np

_rng = np.random.RandomState(1234)


def get_rng():
    return _rng


def set_rng(new_rng):
    global _rng
    _rng = new_rng
