"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference (read-only).

Only `oracle/gen_golden.py` (which writes tests/golden/*.npz) and the optional
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent, e.g. on the GPU box)
use this module.  Nothing in the product package may import it.

The recipe follows SURVEY.md §12: third-party shims on sys.path, `workspace` env var,
argv set BEFORE `lib.config` is imported (argparse runs at import, reference
lib/config/config.py:191-201), CWD = reference root (the `*_path` entries are CWD-relative,
reference lib/config/config.py:166-168).
"""
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("BMV_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_loaded = {}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "networks"))


def load_reference(cfg_file="configs/exps/evaluate/enerf_ours/free_eval.yaml", opts=(), family=None):
    """Import the reference with the given yaml + CLI-style overrides; returns a namespace dict.

    The reference keeps ONE global cfg per process, so a process can hold exactly one
    (cfg_file, opts) combination; a second call with different arguments raises.
    """
    key = (cfg_file, tuple(opts))
    if _loaded:
        if key not in _loaded:
            raise RuntimeError("reference already imported with another cfg in this process")
        return _loaded[key]
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    import torch

    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("workspace", tempfile.mkdtemp(prefix="bmv_ws_"))
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["oracle", "--cfg_file", os.path.join(REFERENCE_ROOT, cfg_file)] + [str(o) for o in opts]
    os.chdir(REFERENCE_ROOT)
    if not torch.cuda.is_available():
        # reference lib/networks/mvsnerf/network.py:44 calls .cuda() in a constructor
        torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        from lib.config import cfg  # noqa
        from lib.networks.enerf import utils as enerf_utils
        from lib.networks.enerf import network as enerf_network
        from lib.datasets import enerf_utils as data_utils
        ns = dict(cfg=cfg, enerf_utils=enerf_utils, enerf_network=enerf_network, data_utils=data_utils)
        if family is None:
            family = "mvsnerf" if "mvsnerf" in cfg_file else "enerf"
        if family == "mvsnerf":
            from lib.networks.mvsnerf import network as mvs_network
            from lib.networks.mvsnerf import utils as mvs_utils
            from lib.networks.mvsnerf import renderer as mvs_renderer
            from lib.networks.boost_mvsnerf import network as boost_mvs_network
            ns.update(mvs_network=mvs_network, mvs_utils=mvs_utils, mvs_renderer=mvs_renderer,
                      boost_mvs_network=boost_mvs_network)
        else:
            from lib.networks.boost_enerf import network as boost_enerf_network
            ns.update(boost_enerf_network=boost_enerf_network)
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    _loaded[key] = ns
    return ns
