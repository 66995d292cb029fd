"""lasagne.init: only the shapes matter here (every value is overwritten by the loaded checkpoint)."""
import numpy as np

from .random import get_rng
from .utils import floatX


class Initializer(object):
    def __call__(self, shape):
        return self.sample(shape)


class Normal(Initializer):
    def __init__(self, std=0.01, mean=0.0):
        self.std, self.mean = std, mean

    def sample(self, shape):
        return floatX(get_rng().normal(self.mean, self.std, size=shape))


class Uniform(Initializer):
    def __init__(self, range=0.01, std=None, mean=0.0):
        self.range = (mean - np.sqrt(3) * std, mean + np.sqrt(3) * std) if std is not None else (
            range if isinstance(range, (tuple, list)) else (-range, range))

    def sample(self, shape):
        return floatX(get_rng().uniform(self.range[0], self.range[1], size=shape))


class Constant(Initializer):
    def __init__(self, val=0.0):
        self.val = val

    def sample(self, shape):
        return floatX(np.ones(shape) * self.val)


class GlorotUniform(Initializer):
    def __init__(self, gain=1.0, c01b=False):
        self.gain = np.sqrt(2) if gain == 'relu' else gain

    def sample(self, shape):
        n1, n2 = shape[:2]
        rf = int(np.prod(shape[2:]))
        a = self.gain * np.sqrt(6.0 / ((n1 + n2) * rf))
        return floatX(get_rng().uniform(-a, a, size=shape))


Glorot = GlorotUniform


class Orthogonal(Initializer):
    def __init__(self, gain=1.0):
        self.gain = np.sqrt(2) if gain == 'relu' else gain

    def sample(self, shape):
        flat = (shape[0], int(np.prod(shape[1:])))
        u, _, v = np.linalg.svd(get_rng().normal(0.0, 1.0, flat), full_matrices=False)
        q = u if u.shape == flat else v
        return floatX(self.gain * q.reshape(shape))
