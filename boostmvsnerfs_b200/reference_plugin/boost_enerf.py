"""Drop-in for reference lib/networks/boost_enerf/network.py — select it from a reference YAML with
    network_module: boostmvsnerfs_b200.reference_plugin.boost_enerf
(the repository root must be on sys.path and reachable from the reference's CWD)."""
import os

from boostmvsnerfs_b200.config import RenderConfig
from boostmvsnerfs_b200.network import BoostEnerfNetwork


def _reference_cfg():
    from lib.config import cfg          # the reference's global yacs config (lib/config/config.py:201)
    return cfg


class Network(BoostEnerfNetwork):
    """Same constructor contract as the reference class (boost_enerf/network.py:11-20):
    `Network()` loads `<cfg.result_dir>/view_selection.json`; `Network(True)` (pre-process) does not."""

    def __init__(self, preprocess=False):
        cfg = _reference_cfg()
        super().__init__(preprocess=preprocess, rc=RenderConfig.from_reference_cfg(cfg),
                         view_selection_file=os.path.join(cfg.result_dir, 'view_selection.json'))
