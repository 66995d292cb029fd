import numpy as np
import theano


def floatX(arr):
    return np.asarray(arr, dtype=theano.config.floatX)


def as_theano_expression(x):
    if isinstance(x, (list, tuple)):
        return [theano.as_var(i) for i in x]
    return theano.as_var(x)


def as_tuple(x, N, t=None):
    try:
        X = tuple(x)
    except TypeError:
        X = (x,) * N
    if t is not None and not all(isinstance(v, (t, np.integer)) for v in X):
        raise TypeError("expected a single value or an iterable of %s, got %r" % (t.__name__, x))
    if len(X) != N:
        raise ValueError("expected a single value or an iterable with length %d, got %r" % (N, x))
    return X


def unique(l):
    seen, out = set(), []
    for el in l:
        if id(el) not in seen:
            seen.add(id(el))
            out.append(el)
    return out


def collect_shared_vars(expressions):
    if isinstance(expressions, theano.Var):
        expressions = [expressions]
    return theano.leaves(expressions)


def create_param(spec, shape, name=None):
    """lasagne.utils.create_param: shared variables and expressions are used as they are (named only if unnamed);
    arrays are wrapped; callables are sampled at `shape`."""
    shape = tuple(shape)
    if any(d is None or d <= 0 for d in shape):
        raise ValueError("Cannot create param with a non-positive shape dimension: %r (%s)" % (shape, name))
    if isinstance(spec, theano.Var):
        if spec.ndim != len(shape):
            raise ValueError("parameter %s: expected %d dimensions, got %d" % (name, len(shape), spec.ndim))
        if not spec.name:
            spec.name = name
        return spec
    if isinstance(spec, np.ndarray):
        if spec.shape != shape:
            raise ValueError("parameter %s: expected shape %r, got %r" % (name, shape, spec.shape))
        return theano.shared(spec, name=name)
    if callable(spec):
        arr = floatX(spec(shape))
        if arr.shape != shape:
            raise ValueError("parameter %s: initializer returned shape %r, expected %r" % (name, arr.shape, shape))
        return theano.shared(arr, name=name)
    raise TypeError("cannot interpret the specification of parameter %s" % name)
