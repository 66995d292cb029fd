"""TEST INFRASTRUCTURE ONLY -- a numpy stand-in for the slice of Lasagne 0.2.dev1 the reference's model files use
(see ../theano/__init__.py for why).  Every class restates the documented behaviour of the Lasagne class of the same
name (lasagne.readthedocs.io, 0.2.dev1: constructor signatures and defaults, parameter names/shapes/tags, output
shapes, `get_output` propagation, `batch_norm` rewiring) on top of the lazy numpy nodes of the theano stand-in."""
from . import random, utils, init, nonlinearities, layers      # noqa: F401

__version__ = "0.2.dev1-shim"
