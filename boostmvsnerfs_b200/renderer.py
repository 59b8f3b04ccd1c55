"""Function-level drop-in: the pure renderer functions of reference lib/networks/enerf/utils.py with
the SAME names, argument order and return shapes, running on the libbmv kernels
(SURVEY.md §8(b) "Function-level boundary").  The only signature difference is the trailing
`rc: RenderConfig`, which replaces the reference's global `cfg`; `patch_reference_utils` binds it
and installs the functions into the reference module so the reference's own Network runs on them.

Batched inputs (leading B) are handled by looping over B; B=1 at test time in the reference.
All tensors must be CUDA float32; there is no CPU path.
"""
import torch

from . import ops
from .config import RenderConfig


def _views(n):
    return list(range(n))


def get_proj_mats(batch, src_scale, tar_scale):
    """reference lib/networks/enerf/utils.py:35-55 — kept as torch ops (tiny; identical arithmetic)."""
    B, S = batch['src_inps'].shape[:2]
    k_src = batch['src_ixts'].clone()
    k_src[:, :, :2] *= src_scale
    p_src = k_src @ batch['src_exts'][:, :, :3]
    k_tar = batch['tar_ixt'].clone()
    k_tar[:, :2] *= tar_scale
    p_tar = k_tar @ batch['tar_ext'][:, :3]
    last = torch.zeros((B, 1, 4), device=p_tar.device, dtype=p_tar.dtype)
    last[:, :, 3] = 1
    inv = torch.inverse(torch.cat((p_tar, last), dim=1))
    return p_src.view(B, S, 3, 4) @ inv.view(B, 1, 4, 4)


def get_depth_values(batch, D, level, device, depth, std, near_far, rc: RenderConfig):
    """reference lib/networks/enerf/utils.py:98-153 -> depth_values (B,D,h,w), near_far (B,2,h,w)."""
    B = len(batch['src_inps'])
    H, W = batch['src_inps'].shape[-2:]
    h, w = int(H * rc.volume_scale[level]), int(W * rc.volume_scale[level])
    planes, nfs = [], []
    for b in range(B):
        if depth is None:
            pl, nf = ops.depth_planes_first(batch['near_far'][b], D, h, w, rc.depth_inv[level])
            pl = pl.view(D, 1, 1).expand(D, h, w).contiguous()
        else:
            if not rc.depth_inv[level - 1]:
                raise NotImplementedError("reference traps here (enerf/utils.py:130)")
            pl, nf = ops.depth_planes_next(depth[b], std[b], near_far[b], D, h, w, rc.depth_inv[level])
        planes.append(pl)
        nfs.append(nf)
    return torch.stack(planes), torch.stack(nfs)


def build_feature_volume(feature, batch, D, depth, std, near_far, level, rc: RenderConfig):
    """reference lib/networks/enerf/utils.py:324-351.  feature (B,S,C,Hs,Ws) ->
    (feature_volume (B,C,D,h,w), depth_values (B,D,h,w), near_far (B,2,h,w))."""
    B, S = feature.shape[:2]
    depth_values, near_far = get_depth_values(batch, D, level, feature.device, depth, std, near_far, rc)
    proj = get_proj_mats(batch, src_scale=rc.im_feat_scale[level], tar_scale=rc.volume_scale[level])
    vols = [ops.cost_volume_var(feature[b], _views(S), proj[b], depth_values[b]) for b in range(B)]
    return torch.stack(vols), depth_values, near_far


def depth_regression(depth_prob, depth_values, level, batch, rc: RenderConfig):
    """reference lib/networks/enerf/utils.py:678-731 (level >= 0 branch) -> depth (B,h,w), std (B,h,w)."""
    out = [ops.depth_regression(depth_prob[b], depth_values[b], rc.depth_inv[level]) for b in range(len(depth_prob))]
    return torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out])


def build_rays(depth, std, batch, training, near_far, level, rc: RenderConfig):
    """reference lib/networks/enerf/utils.py:392-422 -> rays (B,R,12)."""
    rays = batch[f'rays_{level}']
    H0, W0 = batch['src_inps'].shape[-2:]
    H, W = int(H0 * rc.render_scale[level]), int(W0 * rc.render_scale[level])
    cams = _dummy_cams(rays.device)
    out = []
    for b in range(rays.shape[0]):
        o = ops.raygen_sample_fetch(depth[b], std[b], near_far[b], rays[b], H, W, rc.depth_inv[level], 1,
                                    None, None, None, cams, [0], want=("rays12",))
        out.append(o['rays12'])
    return torch.stack(out)


def sample_along_depth(rays, N_samples, level, rc: RenderConfig, H=2, W=2):
    """reference lib/networks/enerf/utils.py:424-443 -> world_xyz (B,R,S,3), uvd (B,R,S,3), z_vals (B,R,S)."""
    outs = [ops.sample_rays12(rays[b], N_samples, rc.depth_inv[level], H, W) for b in range(rays.shape[0])]
    return (torch.stack([o['xyz'] for o in outs]), torch.stack([o['uvd'] for o in outs]),
            torch.stack([o['z_vals'] for o in outs]))


def get_vox_feat(ndc_xyz, feature_volume):
    """reference lib/networks/enerf/utils.py:458-460.  ndc_xyz (B,P,3) in [0,1] -> (B,P,C)."""
    cams = _dummy_cams(ndc_xyz.device)
    out = []
    for b in range(ndc_xyz.shape[0]):
        pts = ndc_xyz[b].contiguous()
        o = ops.fetch_points(pts, pts, 2, 2, feature_volume[b], None, None, cams, [0], want=("vox_feat",))
        out.append(o['vox_feat'])
    return torch.stack(out)


def get_img_feat(xyz, img_feat_rgb, batch, training, level, rc: RenderConfig):
    """reference lib/networks/enerf/utils.py:753-786.  xyz (B,R,S,3), img_feat_rgb (B,V,C+3,H,W)
    (features ++ colours already in [0,1]) -> (B,R*S,V,C+3+4)."""
    B, V, C3, H, W = img_feat_rgb.shape
    out = []
    for b in range(B):
        cams = ops.CameraBlock(batch['src_exts'][b], batch['src_ixts'][b], batch['tar_ext'][b])
        o = ops.fetch_points(xyz[b].reshape(-1, 3), None, H, W, None, img_feat_rgb[b][:, :C3 - 3],
                             img_feat_rgb[b][:, C3 - 3:].contiguous(), cams, _views(V),
                             render_scale=rc.render_scale[level], rgb_affine=(1.0, 0.0), want=("img_feat",))
        out.append(o['img_feat'])
    return torch.stack(out)


def mask_viewport(world_xyz, src_exts, src_ixts, inv_scale):
    """reference lib/networks/enerf/utils.py:510-520.  world_xyz (B,R,S,3) -> (B,R*S,1)."""
    B = world_xyz.shape[0]
    V = src_exts.shape[1]
    inv = inv_scale.detach().cpu()
    out = [ops.mask_viewport(world_xyz[b].reshape(-1, 3), src_exts[b], src_ixts[b], _views(V),
                             (float(inv[b, 0]), float(inv[b, 1]))) for b in range(B)]
    return torch.stack(out).unsqueeze(-1)


def raw2outputs(raw, z_vals, white_bkgd=False):
    """reference lib/networks/enerf/utils.py:605-637.  raw (B,R,S,4), z_vals (B,R,S) or None."""
    res = [ops.composite(raw[b], None if z_vals is None else z_vals[b], white_bkgd) for b in range(raw.shape[0])]
    return {'rgb': torch.stack([r[0] for r in res]),
            'depth': None if z_vals is None else torch.stack([r[1] for r in res]),
            'weights': torch.stack([r[2] for r in res])}


def raw2outputs_blend(raws, masks, z_vals, white_bkgd=False):
    """reference lib/networks/enerf/utils.py:639-667.  raws (B,K,R,S,4); masks (B,K,R,S) as produced by
    merge_mlp_outputs (normalised over K; re-normalising in the kernel is the identity)."""
    if white_bkgd:
        raise NotImplementedError
    B, K = raws.shape[:2]
    masks = masks.view(raws.shape[:4])
    res = [ops.composite_blend(list(raws[b].unbind(0)), list(masks[b].unbind(0)), list(z_vals[b].unbind(0)))
           for b in range(B)]
    return {'rgb': torch.stack([r[0] for r in res]), 'depth': torch.stack([r[1] for r in res]),
            'weights': torch.stack([r[2] for r in res])}


_dummy = {}


def _dummy_cams(device):
    key = str(device)
    if key not in _dummy:
        z = torch.zeros((1, 4, 4), device=device)
        _dummy[key] = ops.CameraBlock(z, torch.zeros((1, 3, 3), device=device),
                                      centers=torch.zeros((1, 3), device=device),
                                      tar_center=torch.zeros(3, device=device))
    return _dummy[key]


PATCHABLE = ("build_feature_volume", "depth_regression", "build_rays", "sample_along_depth", "get_vox_feat",
             "get_img_feat", "mask_viewport", "raw2outputs", "raw2outputs_blend")


def patch_reference_utils(utils_module, rc: RenderConfig = None):
    """Install the kernels into the reference's `lib.networks.enerf.utils` (module attributes are
    looked up at call time).  Returns a dict of the originals for un-patching."""
    if rc is None:
        from lib.config import cfg
        rc = RenderConfig.from_reference_cfg(cfg)
    g = globals()
    originals = {}
    for name in PATCHABLE:
        originals[name] = getattr(utils_module, name)
        fn = g[name]
        if 'rc' in fn.__code__.co_varnames[:fn.__code__.co_argcount]:
            def bound(*a, _fn=fn, **kw):
                return _fn(*a, rc=rc, **kw)
            setattr(utils_module, name, bound)
        else:
            setattr(utils_module, name, fn)
    return originals
