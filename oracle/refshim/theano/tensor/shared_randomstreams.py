"""theano.tensor.shared_randomstreams.RandomStreams stand-in (numpy-RandomState streams; see ../../__init__.py).

Restated behaviour (Theano 0.9 `shared_randomstreams.py` / `raw_random.py`): the stream object owns a seed generator
`RandomState(seed)`; every random variable created through it gets its own `RandomState(seedgen.randint(2**30))`,
advanced each time a function evaluates the variable; `seed(s)` re-seeds the generator and then every existing
stream in creation order.  `shuffle_row_elements(x)` = `permute_row_elements(x, permutation(size=x.shape[:-1],
n=x.shape[-1]))`, and `permutation` fills each row with `random_state.permutation(n)`; out[..., j] = x[..., perm[j]].
"""
import numpy as np

from .. import Var, as_var


class RandomStreams(object):
    def __init__(self, seed=None):
        self.default_instance_seed = seed
        self.gen_seedgen = np.random.RandomState(seed)
        self.state_updates = []
        self._streams = []

    def seed(self, seed=None):
        if seed is None:
            seed = self.default_instance_seed
        self.gen_seedgen.seed(seed)
        for st in self._streams:
            st['rng'] = np.random.RandomState(int(self.gen_seedgen.randint(2 ** 30)))

    def _new_stream(self):
        st = {'rng': np.random.RandomState(int(self.gen_seedgen.randint(2 ** 30)))}
        self._streams.append(st)
        return st

    def permutation(self, size=None, n=1, **kwargs):
        raise NotImplementedError("only shuffle_row_elements is used by the reference")

    def shuffle_row_elements(self, input):
        x = as_var(input)
        st = self._new_stream()

        def run(a):
            out = np.empty_like(a)
            for i in np.ndindex(*a.shape[:-1]):
                out[i] = a[i][st['rng'].permutation(a.shape[-1])]
            return out
        return Var(run, [x], ndim=x.ndim)

    def normal(self, size=None, avg=0.0, std=1.0, **kwargs):
        st = self._new_stream()
        from . import _shape_args
        vs, resolve, n = _shape_args(size)
        return Var(lambda *v: st['rng'].normal(avg, std, resolve(v)), vs, ndim=n)
