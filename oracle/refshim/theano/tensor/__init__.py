"""theano.tensor stand-in (see ../__init__.py): the functions the reference's model files call, as lazy numpy nodes."""
import numpy as np

from .. import Var, InputVar, SharedVariable, as_var, config, _bind   # noqa: F401


def _axes(axis, nd):
    if axis is None:
        return None
    ax = (axis,) if isinstance(axis, (int, np.integer)) else tuple(axis)
    return tuple(a % nd for a in ax)


def _reduce(npf, x, axis):
    x = as_var(x)
    ax = _axes(axis, x.ndim)
    return Var(lambda a: npf(a, axis=ax), [x], ndim=0 if ax is None else x.ndim - len(ax))


def sum(x, axis=None, **kw): return _reduce(np.sum, x, axis)
def mean(x, axis=None, **kw): return _reduce(np.mean, x, axis)
def max(x, axis=None, **kw): return _reduce(np.max, x, axis)
def min(x, axis=None, **kw): return _reduce(np.min, x, axis)
def var(x, axis=None, **kw): return _reduce(np.var, x, axis)
def argmax(x, axis=None, **kw): return _reduce(np.argmax, x, axis)


def cumsum(x, axis=None):
    x = as_var(x)
    return Var(lambda a: np.cumsum(a, axis=axis), [x], ndim=1 if axis is None else x.ndim)


def _unary(npf):
    def f(x):
        x = as_var(x)
        return Var(npf, [x], ndim=x.ndim)
    return f


exp, log, sqrt, tanh, abs_, sgn, neg = map(_unary, (np.exp, np.log, np.sqrt, np.tanh, np.abs, np.sign, np.negative))
sqr = square = _unary(np.square)
expm1, log1p, floor, ceil = map(_unary, (np.expm1, np.log1p, np.floor, np.ceil))
inv = _unary(lambda a: 1.0 / a)
ones_like, zeros_like = _unary(np.ones_like), _unary(np.zeros_like)


def sigmoid(x):
    return _unary(lambda a: 1.0 / (1.0 + np.exp(-a)))(x)


def _binary(npf):
    def f(a, b):
        return as_var(a)._bin(b, npf)
    return f


add, sub, mul, true_div, maximum, minimum, pow = map(
    _binary, (np.add, np.subtract, np.multiply, np.true_divide, np.maximum, np.minimum, np.power))
eq, neq, lt, le, gt, ge = map(_binary, (np.equal, np.not_equal, np.less, np.less_equal, np.greater, np.greater_equal))


def switch(c, a, b):
    c, a, b = as_var(c), as_var(a), as_var(b)
    return Var(np.where, [c, a, b], ndim=builtins_max(c.ndim, a.ndim, b.ndim))


def clip(x, lo, hi):
    x = as_var(x)
    return Var(np.clip, [x, lo, hi], ndim=x.ndim)


import builtins as _b                                            # noqa: E402
builtins_max = _b.max


def dot(a, b):
    a, b = as_var(a), as_var(b)
    return Var(np.dot, [a, b], ndim=builtins_max(a.ndim + b.ndim - 2, 0))


def tensordot(a, b, axes=2):
    a, b = as_var(a), as_var(b)
    n = axes if isinstance(axes, int) else len(axes[0])
    return Var(lambda x, y: np.tensordot(x, y, axes), [a, b], ndim=a.ndim + b.ndim - 2 * n)


def concatenate(tensors, axis=0):
    ts = [as_var(t) for t in tensors]
    return Var(lambda *v: np.concatenate(v, axis=axis), ts, ndim=ts[0].ndim)


def _shape_args(shape):
    """a shape given as a tuple/list whose entries may be ints or scalar Vars (or one 1-d Var)"""
    if isinstance(shape, Var):
        return [shape], lambda vals: tuple(int(s) for s in np.atleast_1d(vals[0])), None
    shape = list(shape) if isinstance(shape, (list, tuple)) else [shape]
    vs = [s for s in shape if isinstance(s, Var)]

    def resolve(vals):
        it = iter(vals)
        return tuple(int(next(it)) if isinstance(s, Var) else int(s) for s in shape)
    return vs, resolve, len(shape)


def reshape(x, shape, ndim=None):
    x = as_var(x)
    vs, resolve, n = _shape_args(shape)
    return Var(lambda a, *vals: a.reshape(resolve(vals)), [x] + vs, ndim=ndim if ndim is not None else n)


def zeros(shape, dtype=None):
    vs, resolve, n = _shape_args(shape)
    return Var(lambda *vals: np.zeros(resolve(vals)), vs, ndim=n)


def ones(shape, dtype=None):
    vs, resolve, n = _shape_args(shape)
    return Var(lambda *vals: np.ones(resolve(vals)), vs, ndim=n)


def eye(n, m=None, k=0, dtype=None):
    n = as_var(n)
    return Var(lambda a: np.eye(int(a)), [n], ndim=2)


def arange(start, stop=None, step=1, dtype=None):
    args = [as_var(a) for a in ((start,) if stop is None else (start, stop))]
    return Var(lambda *v: np.arange(*[float(x) for x in v], step).astype(np.float64), args, ndim=1)


def tile(x, reps, ndim=None):
    x = as_var(x)
    return Var(lambda a: np.tile(a, reps), [x], ndim=builtins_max(x.ndim, len(reps)))


def transpose(x, axes=None):
    x = as_var(x)
    return Var(lambda a: np.transpose(a, axes), [x], ndim=x.ndim)


def split(x, splits_size, n_splits, axis=0):
    x = as_var(x)
    edges = np.cumsum([0] + list(splits_size))
    return [x[(slice(None),) * (axis % x.ndim) + (slice(int(edges[i]), int(edges[i + 1])),)] for i in range(n_splits)]


def nonzero(x, return_matrix=False):
    x = as_var(x)
    return tuple(Var(lambda a, d=d: np.nonzero(a)[d], [x], ndim=1) for d in range(x.ndim))


def _sub_write(sub, y, inc):
    """T.set_subtensor(x[idx], y) / T.inc_subtensor(x[idx], y): a copy of x with x[idx] replaced / incremented"""
    assert hasattr(sub, 'index_of'), "set_subtensor expects an indexing expression x[...]"
    base, idx_leaves = sub.inputs[0], sub.inputs[1:]
    y = as_var(y)

    def run(a, yv, *vals):
        out = np.array(a, dtype=np.result_type(a, yv), copy=True)
        idx = sub.index_of(vals)
        if inc:
            out[idx] += yv
        else:
            out[idx] = yv
        return out
    return Var(run, [base, y] + list(idx_leaves), ndim=base.ndim)


def set_subtensor(sub, y, **kw): return _sub_write(sub, y, False)
def inc_subtensor(sub, y, **kw): return _sub_write(sub, y, True)


# ---- inputs -------------------------------------------------------------------------------------------------
class TensorType(object):
    def __init__(self, dtype, broadcastable):
        self.dtype, self.ndim = dtype, len(broadcastable)

    def __call__(self, name=None):
        return InputVar(name, self.ndim, self.dtype)


def _input(nd):
    def make(name=None, dtype=None):
        return InputVar(name, nd, dtype or config.floatX)
    return make


scalar, vector, matrix, tensor3, tensor4 = map(_input, range(5))
fscalar, fvector, fmatrix, ftensor3, ftensor4 = scalar, vector, matrix, tensor3, tensor4
iscalar = lambda name=None: InputVar(name, 0, 'int32')            # noqa: E731
ivector = lambda name=None: InputVar(name, 1, 'int32')            # noqa: E731


# ---- gradient -------------------------------------------------------------------------------------------------
GRAD_STEP = 1e-6


class _NumericGrad(Var):
    """d cost / d wrt by central differences in float64 (cost re-evaluated 2*size(wrt) times)"""
    def __init__(self, cost, wrt):
        Var.__init__(self, None, (), 'grad', wrt.ndim)
        self.cost, self.wrt = cost, wrt
        self.extra_deps = (cost,)

    def _value(self, memo):
        base = memo['bindings']
        x0 = np.array(memo[id(self.wrt)] if id(self.wrt) in memo else self.wrt._value(memo), dtype=np.float64)
        g = np.zeros_like(x0)
        for i in range(x0.size):
            vals = []
            for sgn in (+1.0, -1.0):
                x = x0.copy()
                x.flat[i] += sgn * GRAD_STEP
                m = dict(base)
                m[id(self.wrt)] = x
                m['bindings'] = dict(m)
                vals.append(float(self.cost._value(m)))
            g.flat[i] = (vals[0] - vals[1]) / (2 * GRAD_STEP)
        return g


def grad(cost, wrt, **kwargs):
    assert as_var(cost).ndim == 0, "grad needs a scalar cost"
    return _NumericGrad(cost, wrt)


from . import shared_randomstreams      # noqa: E402,F401
