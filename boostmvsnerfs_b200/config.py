"""Explicit, immutable view of the `cfg.enerf.*` keys the hot path reads.

The reference threads a global yacs `cfg` through every renderer function
(reference lib/config/config.py:201; keys listed in SURVEY.md §5 "Config / flags",
values from reference configs/exps/pretrain/enerf/dtu_pretrain.yaml:21-47,
configs/exps/pretrain/enerf_ours/dtu_pretrain.yaml:10-15,
configs/exps/evaluate/enerf_ours/base_eval.yaml:3-5,
configs/exps/pretrain/mvsnerf_ours/dtu_pretrain.yaml:12-21).
Our kernels take scalars; the drop-in `Network` reads the reference cfg ONCE through
`RenderConfig.from_reference_cfg` and passes this object down.
"""
from dataclasses import dataclass, field, replace
from typing import Tuple


@dataclass(frozen=True)
class RenderConfig:
    # cfg.enerf.*
    cost_volume_input_views: int = 3
    chunk_size: int = 1000000
    white_bkgd: bool = False
    viewdir_agg: bool = True
    # cfg.enerf.cas_config.*
    num: int = 2
    k_best: int = 4
    depth_inv: Tuple[bool, ...] = (True, False)
    volume_scale: Tuple[float, ...] = (0.125, 0.5)
    volume_planes: Tuple[int, ...] = (64, 8)
    im_feat_scale: Tuple[float, ...] = (0.25, 0.5)
    im_ibr_scale: Tuple[float, ...] = (0.25, 1.0)
    render_scale: Tuple[float, ...] = (0.25, 1.0)
    render_im_feat_level: Tuple[int, ...] = (0, 2)
    nerf_model_feat_ch: Tuple[int, ...] = (32, 8)
    render_if: Tuple[bool, ...] = (False, True)
    num_samples: Tuple[int, ...] = (8, 2)

    @staticmethod
    def enerf_eval(k_best: int = 4) -> "RenderConfig":
        """ENeRF + boost evaluation config (reference configs/exps/evaluate/enerf_ours/*.yaml)."""
        return RenderConfig(k_best=k_best)

    @staticmethod
    def enerf_pretrain(k_best: int = 4) -> "RenderConfig":
        """Both cascade levels rendered (reference configs/exps/pretrain/enerf/dtu_pretrain.yaml:42)."""
        return RenderConfig(k_best=k_best, render_if=(True, True))

    @staticmethod
    def mvsnerf_eval(k_best: int = 4, num_samples: int = 32) -> "RenderConfig":
        """MVSNeRF + boost (reference configs/exps/pretrain/mvsnerf_ours/dtu_pretrain.yaml:12-21):
        one level, uniform planes; the remaining cas_config lists keep the ENeRF parent's values
        but only index 0 is read."""
        return RenderConfig(k_best=k_best, num=1, depth_inv=(False,), render_scale=(1.0,),
                            num_samples=(num_samples,), render_if=(True,))

    @staticmethod
    def from_reference_cfg(cfg) -> "RenderConfig":
        e, c = cfg.enerf, cfg.enerf.cas_config
        return RenderConfig(
            cost_volume_input_views=int(e.get("cost_volume_input_views", 3)),
            chunk_size=int(e.chunk_size), white_bkgd=bool(e.white_bkgd),
            viewdir_agg=bool(e.viewdir_agg), num=int(c.num), k_best=int(c.get("k_best", 1)),
            depth_inv=tuple(bool(x) for x in c.depth_inv),
            volume_scale=tuple(float(x) for x in c.volume_scale),
            volume_planes=tuple(int(x) for x in c.volume_planes),
            im_feat_scale=tuple(float(x) for x in c.im_feat_scale),
            im_ibr_scale=tuple(float(x) for x in c.im_ibr_scale),
            render_scale=tuple(float(x) for x in c.render_scale),
            render_im_feat_level=tuple(int(x) for x in c.render_im_feat_level),
            nerf_model_feat_ch=tuple(int(x) for x in c.nerf_model_feat_ch),
            render_if=tuple(bool(x) for x in c.render_if),
            num_samples=tuple(int(x) for x in c.num_samples))

    def with_(self, **kw) -> "RenderConfig":
        return replace(self, **kw)
