"""Shim for the stdlib `imp` module (removed in Python 3.12); the reference's plugin factory
uses `imp.load_source` (reference lib/networks/make_network.py:1-10). Test infrastructure only."""
import importlib.util
import sys


def load_source(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module
