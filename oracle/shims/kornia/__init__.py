"""Test-infrastructure shim for the un-vendored `kornia` dependency of the reference.

Only `kornia.utils.create_meshgrid` is used on the hot path
(reference lib/networks/enerf/utils.py:4,65 and lib/networks/mvsnerf/utils.py:578,603).
This shim exists so `oracle/gen_golden.py` can import the UNMODIFIED reference in a
container that has no kornia wheel.  It is never imported by the product package.
"""
from . import utils  # noqa: F401
