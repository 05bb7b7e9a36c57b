"""MRG_RandomStreams stand-in (test infrastructure).  Theano's MRG31k3p stream is not reproduced: the fixtures are
generated with noise = 0, where no sample is ever drawn; drawing one raises."""


class MRG_RandomStreams(object):
    def __init__(self, seed=None, **kwargs):
        self.seed = seed

    def normal(self, *args, **kwargs):
        raise NotImplementedError('random streams are not reproduced by the stand-in; use noise = 0 / deterministic graphs')

    uniform = binomial = normal
