"""theano.tensor.opt stand-in: graph-optimizer registration is meaningless for eager numpy evaluation."""


def register_canonicalize(fn=None, *a, **k):
    return fn
