"""Plugin files for the reference's network factory.

The reference instantiates its network with
    imp.load_source(cfg.network_module, cfg.network_path).Network([preprocess])
(reference lib/networks/make_network.py:3-10), where `network_path` is `network_module` with dots
replaced by slashes plus ".py", relative to the CWD (reference lib/config/config.py:166-168).
`boost_enerf.py` / `enerf.py` in this directory are such files: each exposes `class Network`.
See INTEGRATION.md for the two-line YAML change on the reference side.
"""
