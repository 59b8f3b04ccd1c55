"""Shim: the reference drops into `ipdb.set_trace()` on NaN / out-of-range guards
(e.g. reference lib/networks/enerf/utils.py:93-94). Under test that must be a hard error."""


def set_trace(*args, **kwargs):
    raise RuntimeError("reference hit an ipdb.set_trace() guard")
