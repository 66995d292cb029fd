"""`imp.load_source` for the reference's API.py (the `imp` module left the standard library in Python 3.12)."""
import importlib.util
import sys


def load_source(name, pathname):
    spec = importlib.util.spec_from_file_location(name, pathname)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module
