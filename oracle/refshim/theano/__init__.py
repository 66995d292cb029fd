"""TEST INFRASTRUCTURE ONLY -- a numpy stand-in for the slice of Theano the reference's model files touch.

Why: the reference (ajbrock/Neural-Photo-Editor) is Python over Theano 0.9 / Lasagne 0.2.dev1, neither of which is
installed (or installable) here, so its files cannot be imported as they are.  With this package and `../lasagne`
first on sys.path, the reference's OWN, UNMODIFIED `API.py`, `IAN_simple.py`, `IANv1.py`, `IAN.py`, `layers.py`,
`mask_generator.py` and `GANcheckpoints.py` execute from /root/reference: their layer wiring, hyper-parameters,
parameter names, weight loading and function construction are the reference's code, not a reading of it.  What is
restated here is only the third-party semantics underneath (documented Theano/Lasagne/cuDNN behaviour), in float64.
`tests/golden/make_golden_ref.py` uses it to generate `tests/golden/ref_exec_*.npz`; nothing in the product or in
bench.py imports it.

Model: every expression is a lazy node (`Var`) holding a numpy function of its inputs; shared variables are leaves
whose current value is read at evaluation time (so `set_value` after graph construction works, as in Theano);
`theano.function` binds inputs and evaluates.  `tensor.grad` differentiates numerically (central differences in
float64) -- deliberately independent of the analytic backward pass the oracle and the CUDA kernels implement.
"""
import numpy as np


class _Config(object):
    floatX = 'float64'          # evaluate the reference graph in double: the fixtures are definitional


config = _Config()


def _is_float(a):
    return np.issubdtype(np.asarray(a).dtype, np.floating)


class Var(object):
    """lazy expression node: value = fn(*[value of each input])"""
    def __init__(self, fn, inputs=(), name=None, ndim=None):
        self._fn, self.inputs, self.name, self._ndim = fn, tuple(inputs), name, ndim

    # ---- evaluation
    def _value(self, memo):
        k = id(self)
        if k not in memo:
            memo[k] = self._fn(*[a._value(memo) if isinstance(a, Var) else a for a in self.inputs])
        return memo[k]

    def eval(self, givens=None):
        return self._value(_bind(givens or {}))

    # ---- static rank (DenseLayer & co. branch on input.ndim while the graph is built)
    @property
    def ndim(self):
        if self._ndim is None:
            dummy = [np.zeros((1,) * a.ndim) if isinstance(a, Var) else a for a in self.inputs]
            self._ndim = int(np.ndim(self._fn(*dummy)))
        return self._ndim

    @property
    def shape(self):
        return Var(lambda a: np.asarray(a.shape, dtype=np.int64), [self], ndim=1)

    @property
    def dtype(self):
        return config.floatX

    @property
    def T(self):
        return Var(lambda a: a.T, [self], ndim=self.ndim)

    # ---- arithmetic
    def _bin(self, other, f, swap=False):
        a, b = (other, self) if swap else (self, other)
        nd = max(np.ndim(x) if not isinstance(x, Var) else x.ndim for x in (a, b))
        return Var(f, [a, b], ndim=nd)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._bin(o, np.true_divide, True)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __pow__(self, o): return self._bin(o, np.power)
    def __neg__(self): return Var(np.negative, [self], ndim=self.ndim)
    def __abs__(self): return Var(np.abs, [self], ndim=self.ndim)
    def __lt__(self, o): return self._bin(o, np.less)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    __hash__ = object.__hash__            # `==` stays identity, as for Theano variables

    # ---- indexing: ints, slices, None, Ellipsis; slice bounds and indices may themselves be Vars
    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        leaves = []

        def strip(x):
            if isinstance(x, Var):
                leaves.append(x)
                return ('var', len(leaves) - 1)
            if isinstance(x, slice):
                return slice(strip(x.start), strip(x.stop), strip(x.step))
            return x
        spec = tuple(strip(i) for i in idx)

        def build(x, vals):
            if isinstance(x, tuple) and len(x) == 2 and x[0] == 'var':
                v = np.asarray(vals[x[1]])
                return int(v) if v.ndim == 0 else v
            if isinstance(x, slice):
                return slice(build(x.start, vals), build(x.stop, vals), build(x.step, vals))
            return x

        def run(a, *vals):
            return a[tuple(build(i, vals) for i in spec)]
        # rank: apply the index to a dummy with scalar Vars -> 0
        dummy = np.zeros((4,) * self.ndim)[tuple(build(i, [0] * len(leaves)) for i in spec)]
        out = Var(run, [self] + leaves, ndim=dummy.ndim)
        out.index_of = lambda vals: tuple(build(i, vals) for i in spec)      # for set_subtensor / inc_subtensor
        return out

    # ---- tensor methods the reference calls
    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])
        kept = [p for p in pattern if p != 'x']

        def run(a):
            drop = [ax for ax in range(a.ndim) if ax not in kept]
            assert all(a.shape[ax] == 1 for ax in drop), "dimshuffle can only drop broadcastable axes"
            a = a.transpose(kept + drop).reshape([a.shape[ax] for ax in kept])
            shape, it = [], iter(a.shape)
            for p in pattern:
                shape.append(1 if p == 'x' else next(it))
            return a.reshape(shape)
        return Var(run, [self], ndim=len(pattern))

    def flatten(self, ndim=1):
        return Var(lambda a: a.reshape(a.shape[:ndim - 1] + (-1,)), [self], ndim=ndim)

    def reshape(self, shape, ndim=None):
        from . import tensor
        return tensor.reshape(self, shape, ndim)

    def astype(self, dtype):
        return Var(lambda a: a.astype(config.floatX if dtype == 'floatX' else dtype), [self], ndim=self.ndim)

    def sum(self, axis=None): from . import tensor; return tensor.sum(self, axis)
    def mean(self, axis=None): from . import tensor; return tensor.mean(self, axis)
    def max(self, axis=None): from . import tensor; return tensor.max(self, axis)
    def min(self, axis=None): from . import tensor; return tensor.min(self, axis)
    def var(self, axis=None): from . import tensor; return tensor.var(self, axis)

    def __repr__(self):
        return "<%s %s ndim=%s>" % (type(self).__name__, self.name, self._ndim)


Variable = Var


class InputVar(Var):
    """a function argument (T.TensorType(...)('X'), T.scalar(...), ...)"""
    def __init__(self, name, ndim, dtype):
        Var.__init__(self, None, (), name, ndim)
        self.in_dtype = dtype

    def _value(self, memo):
        if id(self) not in memo:
            raise ValueError("no value bound to input %r" % (self.name,))
        return memo[id(self)]


class SharedVariable(Var):
    def __init__(self, value, name=None):
        value = np.array(value)
        Var.__init__(self, None, (), name, value.ndim)
        self.value = value

    def _value(self, memo):
        v = memo.get(id(self), self.value)          # stored as given (float32 checkpoints); computed in double
        return v.astype(np.float64) if _is_float(v) else v

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, v, borrow=False):
        self.value = np.array(v)


def shared(value, name=None, borrow=False, **kwargs):
    return SharedVariable(value, name)


def _as_value(var, val):
    val = np.asarray(val)
    dt = getattr(var, 'in_dtype', None)
    if dt and ('int' in dt):
        return val.astype(np.int64)
    return val.astype(np.float64) if _is_float(val) else val


def _bind(givens):
    memo = {id(v): _as_value(v, x) for v, x in givens.items()}
    memo['bindings'] = dict(memo)
    return memo


def as_var(x):
    if isinstance(x, Var):
        return x
    c = np.asarray(x)
    return Var(lambda: c, (), ndim=c.ndim)


def leaves(exprs, kind=SharedVariable):
    """shared variables an expression depends on, depth-first, unique (theano.gof.graph.inputs order)"""
    seen, out = set(), []

    def walk(v):
        if not isinstance(v, Var) or id(v) in seen:
            return
        seen.add(id(v))
        if isinstance(v, kind):
            out.append(v)
        for a in v.inputs:
            walk(a)
        for a in getattr(v, 'extra_deps', ()):
            walk(a)
    for e in exprs:
        walk(e)
    return out


class _Function(object):
    def __init__(self, inputs, outputs=None, updates=None, name=None, givens=None, **kwargs):
        self.inputs, self.outputs, self.name = list(inputs), outputs, name
        self.updates = list(updates.items()) if isinstance(updates, dict) else list(updates or [])

    def __call__(self, *args):
        assert len(args) == len(self.inputs), "%s: expected %d arguments" % (self.name, len(self.inputs))
        memo = _bind(dict(zip(self.inputs, args)))
        outs = self.outputs
        if isinstance(outs, (list, tuple)):
            res = [o._value(memo) for o in outs]
        else:
            res = None if outs is None else outs._value(memo)
        new = [(s, np.array(e._value(memo) if isinstance(e, Var) else e)) for s, e in self.updates]
        for s, v in new:                     # all updates are computed from the old state, then applied
            s.set_value(v)
        return res


def function(inputs, outputs=None, updates=None, name=None, **kwargs):
    return _Function(inputs, outputs, updates, name, **kwargs)


def clone(output, replace=None, **kwargs):
    raise NotImplementedError("theano.clone is only used by the reference's training scripts")


from . import tensor            # noqa: E402  (theano.tensor must be importable as an attribute)
