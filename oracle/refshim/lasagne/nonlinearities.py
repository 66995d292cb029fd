"""lasagne.nonlinearities on lazy numpy nodes (definitions as documented: rectify = max(0,x), LeakyRectify(a) =
max(x, a*x) for 0<=a<=1, elu = x if x>0 else exp(x)-1, sigmoid, tanh, softmax over axis 1, linear = identity)."""
import numpy as np
import theano.tensor as T


def _elem(f):
    return lambda x: T._unary(f)(x)


sigmoid = _elem(lambda a: 1.0 / (1.0 + np.exp(-a)))
tanh = _elem(np.tanh)
rectify = _elem(lambda a: np.where(a > 0, a, 0.0 * a))
elu = _elem(lambda a: np.where(a > 0, a, np.expm1(np.minimum(a, 0.0))))
softplus = _elem(lambda a: np.log1p(np.exp(a)))
softmax = _elem(lambda a: np.exp(a - a.max(axis=1, keepdims=True)) / np.exp(a - a.max(axis=1, keepdims=True)).sum(axis=1, keepdims=True))


def linear(x):
    return x


identity = linear


class LeakyRectify(object):
    def __init__(self, leakiness=0.01):
        self.leakiness = leakiness

    def __call__(self, x):
        k = self.leakiness
        return T._unary(lambda a: np.where(a > 0, a, k * a))(x)


leaky_rectify = LeakyRectify(0.01)
very_leaky_rectify = LeakyRectify(1. / 3)
