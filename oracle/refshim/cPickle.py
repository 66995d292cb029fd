"""Python 2's cPickle, for the reference's GANcheckpoints.py."""
from pickle import *            # noqa: F401,F403
from pickle import dumps, loads, dump, load   # noqa: F401
