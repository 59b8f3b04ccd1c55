"""Drop-in for reference lib/networks/boost_mvsnerf/network.py — YAML:
    network_module: boostmvsnerfs_b200.reference_plugin.boost_mvsnerf"""
import os

from boostmvsnerfs_b200.config import RenderConfig
from boostmvsnerfs_b200.network_mvs import BoostMvsnerfNetwork


class Network(BoostMvsnerfNetwork):
    def __init__(self, preprocess=False):
        from lib.config import cfg
        super().__init__(preprocess=preprocess, rc=RenderConfig.from_reference_cfg(cfg),
                         view_selection_file=os.path.join(cfg.result_dir, 'view_selection.json'))
