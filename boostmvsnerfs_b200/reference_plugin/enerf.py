"""Drop-in for reference lib/networks/enerf/network.py (single-volume ENeRF baseline) — YAML:
    network_module: boostmvsnerfs_b200.reference_plugin.enerf"""
from boostmvsnerfs_b200.config import RenderConfig
from boostmvsnerfs_b200.network import EnerfNetwork


class Network(EnerfNetwork):
    def __init__(self):
        from lib.config import cfg
        super().__init__(rc=RenderConfig.from_reference_cfg(cfg))
