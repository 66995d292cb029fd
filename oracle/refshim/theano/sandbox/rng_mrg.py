"""theano.sandbox.rng_mrg.MRG_RandomStreams stand-in (see ../__init__.py).

The MRG31k3p generator itself is NOT emulated.  The reference's inference path needs it for exactly one thing: the
MADE mask generator draws each hidden unit's connectivity from `multinomial(pvals)` with `l = mask_distribution = 0`
(reference layers.py:758, mask_generator.py:66-70), which makes every `pvals` row one-hot -- the draw is then the same
for any generator.  `multinomial` therefore insists on one-hot rows; `normal` (GaussianSampleLayer's noise, unused with
deterministic=True) draws from numpy.
"""
import numpy as np

from .. import Var, as_var
from ..tensor import _shape_args


class MRG_RandomStreams(object):
    def __init__(self, seed=12345, use_cuda=None):
        self.rstate = np.asarray([seed] * 6 if isinstance(seed, int) else seed, dtype=np.int64)
        self.state_updates = []
        self._np = np.random.RandomState(int(self.rstate[0]) % (2 ** 31))

    def multinomial(self, size=None, n=1, pvals=None, ndim=None, dtype='int64', nstreams=None):
        p = as_var(pvals)

        def run(pv):
            pv = np.asarray(pv, np.float64)
            pv = pv / pv.sum(axis=1, keepdims=True)
            if not np.all((pv == 0) | (pv == 1)):
                raise NotImplementedError("MRG31k3p is not emulated: multinomial only accepts one-hot rows here")
            return pv.copy()
        return Var(run, [p], ndim=2)

    def normal(self, size=None, avg=0.0, std=1.0, ndim=None, dtype=None, nstreams=None):
        vs, resolve, n = _shape_args(size)
        return Var(lambda *v: self._np.normal(avg, std, resolve(v)), vs, ndim=n)

    def uniform(self, size=None, low=0.0, high=1.0, ndim=None, dtype=None, nstreams=None):
        vs, resolve, n = _shape_args(size)
        return Var(lambda *v: self._np.uniform(low, high, resolve(v)), vs, ndim=n)
