"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container (the GPU box has no /root/reference):

    python oracle/gen_golden.py ops            # op-level vectors   -> tests/golden/enerf_ops.npz
    python oracle/gen_golden.py chain_eval     # Network.forward    -> tests/golden/enerf_chain_eval.npz
    python oracle/gen_golden.py chain_pretrain # both levels render -> tests/golden/enerf_chain_pretrain.npz
    python oracle/gen_golden.py single         # enerf.Network      -> tests/golden/enerf_single.npz
    python oracle/gen_golden.py mvs_ops        # MVSNeRF op-level   -> tests/golden/mvsnerf_ops.npz
    python oracle/gen_golden.py mvs_chain      # boost_mvsnerf.Network.forward -> tests/golden/mvsnerf_chain.npz
    python oracle/gen_golden.py mvs_chain_d32  # same at 32 / 128 depth planes -> tests/golden/mvsnerf_chain_d{32,128}.npz
    python oracle/gen_golden.py mvs_chain_d128
    python oracle/gen_golden.py all            # each of the above in its own process

Every stored array is either an INPUT handed to a reference function or the OUTPUT the reference
returned for it; nothing here is computed by this repository's oracle or kernels.
"""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

from boostmvsnerfs_b200.synth import make_scene  # noqa: E402  (input generator only)
from oracle.ref_loader import load_reference  # noqa: E402

TINY = dict(H=64, W=96, n_views=4)


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, arrays):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrays)} arrays")


def gen_ops():
    ns = load_reference("configs/exps/evaluate/enerf_ours/free_eval.yaml", ["enerf.cas_config.k_best", 2])
    U, cfg = ns["enerf_utils"], ns["cfg"]
    cc = cfg.enerf.cas_config
    g = torch.Generator().manual_seed(1234)
    out = {}
    scene = make_scene(seed=3, smooth=True, **TINY)
    H, W = TINY["H"], TINY["W"]
    triple = [0, 2, 3]
    batch = dict(scene)
    batch["src_inps"] = scene["all_src_inps"][:, triple]
    batch["src_exts"] = scene["all_src_exts"][:, triple]
    batch["src_ixts"] = scene["all_src_ixts"][:, triple]
    for k in ("src_inps", "src_exts", "src_ixts", "tar_ext", "tar_ixt", "near_far", "rays_0", "rays_1"):
        out[f"in_{k}"] = _np(batch[k])

    # ---- level 0: planes, projection matrices, cost volume, depth regression
    feat0 = torch.randn(1, 3, 32, H // 4, W // 4, generator=g)
    feat1 = torch.randn(1, 3, 16, H // 2, W // 2, generator=g)
    out["in_feat0"], out["in_feat1"] = _np(feat0), _np(feat1)
    pm0 = U.get_proj_mats(batch, src_scale=cc.im_feat_scale[0], tar_scale=cc.volume_scale[0])
    out["proj_mats_l0"] = _np(pm0)
    vol0, planes0, nf0 = U.build_feature_volume(feat0, batch, D=cc.volume_planes[0], depth=None, std=None,
                                                near_far=None, level=0)
    out["planes_l0"], out["near_far_l0"], out["volume_l0"] = _np(planes0), _np(nf0), _np(vol0)
    warped, _ = U.homo_warp(feat0[:, 1], pm0[:, 1], planes0)
    out["warped_l0_view1"] = _np(warped)
    logits0 = torch.randn(1, cc.volume_planes[0], H // 8, W // 8, generator=g) * 2
    out["in_logits0"] = _np(logits0)
    depth0, std0 = U.depth_regression(logits0.clone(), planes0, 0, batch)
    out["depth_l0"], out["std_l0"] = _np(depth0), _np(std0)

    # ---- level 1
    pm1 = U.get_proj_mats(batch, src_scale=cc.im_feat_scale[1], tar_scale=cc.volume_scale[1])
    out["proj_mats_l1"] = _np(pm1)
    vol1, planes1, nf1 = U.build_feature_volume(feat1, batch, D=cc.volume_planes[1], depth=depth0, std=std0,
                                                near_far=nf0, level=1)
    out["planes_l1"], out["near_far_l1"], out["volume_l1"] = _np(planes1), _np(nf1), _np(vol1)
    logits1 = torch.randn(1, cc.volume_planes[1], H // 2, W // 2, generator=g) * 2
    out["in_logits1"] = _np(logits1)
    depth1, std1 = U.depth_regression(logits1.clone(), planes1, 1, batch)
    out["depth_l1"], out["std_l1"] = _np(depth1), _np(std1)

    # ---- rays / samples, level 1 (depth_inv False) and level 0 (depth_inv True)
    rays12_l1 = U.build_rays(depth1, std1, batch, False, nf1, 1)
    out["rays12_l1"] = _np(rays12_l1)
    xyz1, uvd1, z1 = U.sample_along_depth(rays12_l1, N_samples=cc.num_samples[1], level=1)
    out["xyz_l1"], out["uvd_l1"], out["z_l1"] = _np(xyz1), _np(uvd1), _np(z1)
    xyz1s, uvd1s, z1s = U.sample_along_depth(rays12_l1, N_samples=1, level=1)
    out["xyz_l1_s1"], out["uvd_l1_s1"], out["z_l1_s1"] = _np(xyz1s), _np(uvd1s), _np(z1s)
    rays12_l0 = U.build_rays(depth0, std0, batch, False, nf0, 0)
    out["rays12_l0"] = _np(rays12_l0)
    xyz0, uvd0, z0 = U.sample_along_depth(rays12_l0, N_samples=cc.num_samples[0], level=0)
    out["xyz_l0"], out["uvd_l0"], out["z_l0"] = _np(xyz0), _np(uvd0), _np(z0)

    # ---- fetches at level 1
    regvol1 = torch.randn(1, 8, cc.volume_planes[1], H // 2, W // 2, generator=g)
    out["in_regvol1"] = _np(regvol1)
    uvdn = uvd1.clone()
    uvdn[..., 0], uvdn[..., 1] = uvdn[..., 0] / (W - 1), uvdn[..., 1] / (H - 1)
    out["vox_feat_l1"] = _np(U.get_vox_feat(uvdn.reshape(1, -1, 3), regvol1))
    imfeat2 = torch.randn(1, 3, 8, H, W, generator=g)
    out["in_imfeat2"] = _np(imfeat2)
    rgbs = U.unpreprocess(batch["src_inps"], render_scale=cc.render_scale[1])
    out["unpreprocess_l1"] = _np(rgbs)
    out["img_feat_l1"] = _np(U.get_img_feat(xyz1, torch.cat((imfeat2, rgbs), dim=2), batch, False, 1))
    # ---- fetches at level 0 (render scale 0.25: resize + scaled intrinsics, 32+3 channels)
    regvol0 = torch.randn(1, 8, cc.volume_planes[0], H // 8, W // 8, generator=g)
    out["in_regvol0"] = _np(regvol0)
    H0, W0 = int(H * cc.render_scale[0]), int(W * cc.render_scale[0])
    uvdn0 = uvd0.clone()
    uvdn0[..., 0], uvdn0[..., 1] = uvdn0[..., 0] / (W0 - 1), uvdn0[..., 1] / (H0 - 1)
    out["vox_feat_l0"] = _np(U.get_vox_feat(uvdn0.reshape(1, -1, 3), regvol0))
    rgbs0 = U.unpreprocess(batch["src_inps"], render_scale=cc.render_scale[0])
    out["unpreprocess_l0"] = _np(rgbs0)
    out["img_feat_l0"] = _np(U.get_img_feat(xyz0, torch.cat((feat0, rgbs0), dim=2), batch, False, 0))

    # ---- 3-D visibility: the tiny rig sees everything, so widen the sample cloud to hit all counts
    inv_scale = torch.tensor([[W - 1, H - 1]], dtype=torch.float32)
    out["mask_l1"] = _np(U.mask_viewport(xyz1, batch["src_exts"], batch["src_ixts"], inv_scale))
    wide = (torch.rand(1, 4096, 2, 3, generator=g) - 0.5) * torch.tensor([14.0, 10.0, 24.0])
    out["in_xyz_wide"] = _np(wide)
    out["mask_wide"] = _np(U.mask_viewport(wide, batch["src_exts"], batch["src_ixts"], inv_scale))
    out["ndc_wide_view0"] = _np(U.get_ndc_coords(wide, batch["src_exts"][:, 0], batch["src_ixts"][:, 0], inv_scale))

    # ---- compositing: single volume and K-blend (merge normalisation restated from
    #      reference lib/networks/boost_enerf/network.py:163-170 via the Network method itself)
    K, R, S = 3, 2048, 2
    raws = torch.randn(1, K, R, S, 4, generator=g)
    raws[..., 3] = torch.nn.functional.softplus(raws[..., 3] * 3)
    raws[..., :3] = torch.sigmoid(raws[..., :3])
    masks = torch.randint(0, 4, (1, K, R, S), generator=g).float() / 3
    masks[:, :, :64] = 0                     # rays nobody sees -> 1/K branch
    zs = torch.rand(1, K, R, S, generator=g) * 6 + 2
    out["in_blend_raws"], out["in_blend_masks"], out["in_blend_z"] = _np(raws), _np(masks), _np(zs)
    net = ns["boost_enerf_network"].Network(preprocess=True)
    merged = net.merge_mlp_outputs(
        {**{f"net_output_view{k}": raws[:, k] for k in range(K)},
         **{f"mask_view{k}": masks[:, k] for k in range(K)},
         **{f"z_vals_view{k}": zs[:, k] for k in range(K)}}, K)
    for key in ("rgb", "depth", "weights"):
        out[f"blend_{key}"] = _np(merged[key])
    S8 = 8
    raw8 = torch.randn(1, R, S8, 4, generator=g)
    raw8[..., 3] = torch.nn.functional.softplus(raw8[..., 3] * 3)
    z8 = torch.sort(torch.rand(1, R, S8, generator=g) * 6 + 2, dim=-1).values
    out["in_comp_raw"], out["in_comp_z"] = _np(raw8), _np(z8)
    single = U.raw2outputs(raw8, z8, False)
    for key in ("rgb", "depth", "weights"):
        out[f"comp_{key}"] = _np(single[key])
    _save("enerf_ops.npz", out)


def _chain(case):
    opts = ["enerf.cas_config.k_best", 2]
    if case == "chain_pretrain":
        opts += ["enerf.cas_config.render_if", "[True,True]"]
    ns = load_reference("configs/exps/evaluate/enerf_ours/free_eval.yaml", opts)
    torch.manual_seed(7)
    net = ns["boost_enerf_network"].Network(preprocess=True).eval()
    # non-trivial BN statistics so the CNNs are not an identity-ish normalisation
    g = torch.Generator().manual_seed(11)
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
    scene = make_scene(seed=5, smooth=True, **TINY)
    k_best = [3, 0]
    net.view_selection_outputs = {"synth_0": k_best}
    batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()}
    with torch.no_grad():
        ret = net(batch)
        # view-selection pre-process of the reference (boost_enerf/network.py:22-121) on the same scene
        vs_batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()}
        sel = net.forward_view_selection(vs_batch)
        cm = net.calc_mask(torch.tensor([0, 1, 3]), {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()})
    out = {f"out_{k}": _np(v) for k, v in ret.items()}
    out["view_selection"] = np.array(sel["synth_0"], dtype=np.int64)
    for k, v in cm.items():
        out[f"calc_mask_013_{k}"] = _np(v)
    for k in ("all_src_inps", "all_src_exts", "all_src_ixts", "tar_ext", "tar_ixt", "near_far", "rays_0", "rays_1"):
        out[f"in_{k}"] = _np(scene[k])
    out["k_best"] = np.array(k_best, dtype=np.int64)
    for k in ("src_inps", "src_exts", "src_ixts"):
        out[f"after_{k}"] = _np(batch[k])
    for name, t in net.state_dict().items():
        out[f"sd_{name}"] = _np(t)
    _save(f"enerf_{case}.npz", out)


def gen_single():
    ns = load_reference("configs/exps/evaluate/enerf/free_eval.yaml", [])
    torch.manual_seed(9)
    net = ns["enerf_network"].Network().eval()
    scene = make_scene(seed=6, smooth=True, H=64, W=96, n_views=3)
    batch = dict(scene)
    batch["src_inps"], batch["src_exts"], batch["src_ixts"] = (
        scene["all_src_inps"], scene["all_src_exts"], scene["all_src_ixts"])
    with torch.no_grad():
        ret = net(batch)
    out = {f"out_{k}": _np(v) for k, v in ret.items()}
    for k in ("all_src_inps", "all_src_exts", "all_src_ixts", "tar_ext", "tar_ixt", "near_far", "rays_0", "rays_1"):
        out[f"in_{k}"] = _np(scene[k])
    out["render_if"] = np.array(list(ns["cfg"].enerf.cas_config.render_if))
    for name, t in net.state_dict().items():
        out[f"sd_{name}"] = _np(t)
    _save("enerf_single.npz", out)


MVS_CFG = "configs/exps/evaluate/mvsnerf_ours/free_eval.yaml"


class _ZeroedEmpty:
    """The reference allocates the 41-channel volume with torch.empty and never writes channels 0:3 of
    the 24-px border (mvsnerf/network.py:906) -> nondeterministic bytes.  While the reference runs we
    make torch.empty return zero-filled memory (an allocator property, not a change to the reference),
    which is the value the oracle and the kernels define for those bytes."""

    def __enter__(self):
        self._orig = torch.empty
        torch.empty = lambda *a, **k: self._orig(*a, **k).zero_()

    def __exit__(self, *a):
        torch.empty = self._orig


def _mvs_scene(seed):
    return make_scene(seed=seed, smooth=True, mvs_near_far_cols=True, render_scales=(1.0,), **TINY)


def gen_mvs_ops():
    ns = load_reference(MVS_CFG, ["enerf.cas_config.k_best", 2, "enerf.cas_config.num_samples", "[8]"])
    MU, MN, BN_, EU = ns["mvs_utils"], ns["mvs_network"], ns["boost_mvs_network"], ns["enerf_utils"]
    g = torch.Generator().manual_seed(4321)
    out = {}
    scene = _mvs_scene(8)
    H, W = TINY["H"], TINY["W"]
    h, w, D, S = H // 4, W // 4, 8, 8
    triple = [1, 0, 3]
    batch = dict(scene)
    batch["src_inps"] = scene["all_src_inps"][:, triple]
    batch["src_exts"] = scene["all_src_exts"][:, triple]
    batch["src_ixts"] = scene["all_src_ixts"][:, triple]
    for k in ("src_inps", "src_exts", "src_ixts", "rays_0", "depth_ranges"):
        out[f"in_{k}"] = _np(batch[k])
    torch.manual_seed(1)
    net = BN_.Network(preprocess=True).eval()
    pm = net.get_proj_mats(batch)
    out["proj_mats"] = _np(pm)
    feats = torch.randn(1, 3, 32, h, w, generator=g)
    out["in_feats"] = _np(feats)
    t = torch.linspace(0., 1., steps=D)
    near, far = batch["depth_ranges"][:, triple].min() * 0.8, batch["depth_ranges"][:, triple].max() * 1.2
    planes = (near * (1. - t) + far * t).unsqueeze(0)
    out["planes"] = _np(planes)
    out["near_far"] = _np(torch.stack([near, far]))
    with _ZeroedEmpty():
        vol = net.build_volume_costvar_img(batch["src_inps"], feats, pm, planes, pad=24)
    out["volume41"] = _np(vol)
    # ---- marching, NDC, fetches, MLP input (every 8th ray keeps the fixture small)
    rays = batch["rays_0"][:, ::8].contiguous()
    out["in_rays_sub"] = _np(rays)
    xyz, z = net.ray_marcher(rays, S)
    out["march_xyz"], out["march_z"] = _np(xyz), _np(z)
    inv_scale = torch.tensor([W - 1, H - 1], dtype=torch.float32)
    ndc = MU.get_ndc_coordinate(batch["src_exts"][0][0], batch["src_ixts"][0][0], xyz[0], inv_scale,
                                near=near, far=far, pad=24)[None]
    out["ndc"] = _np(ndc)
    regvol = torch.randn(1, 8, D, h + 48, w + 48, generator=g)
    out["in_regvol"] = _np(regvol)
    batch["near_far"] = torch.stack([near, far])
    raw_in = net.rendering(batch, xyz[0], ndc, z, rays[..., :3], rays[..., 3:6], regvol)
    out["mlp_input"] = _np(raw_in)
    out["mask"] = _np(EU.mask_viewport(xyz, batch["src_exts"], batch["src_ixts"], inv_scale))
    with torch.no_grad():
        out["mlp_output"] = _np(net.nerf(raw_in))
        for name, tns in net.nerf.state_dict().items():
            out[f"sd_nerf.{name}"] = _np(tns)
    # ---- view-selection coverage mask (no network involved, boost_mvsnerf/network.py:23-45)
    m = net.calc_mask(torch.tensor(triple), dict(scene))
    out["calc_mask"] = _np(m["mask_level0"])
    _save("mvsnerf_ops.npz", out)


def gen_mvs_chain():
    ns = load_reference(MVS_CFG, ["enerf.cas_config.k_best", 2, "enerf.cas_config.num_samples", "[8]"])
    torch.manual_seed(13)
    net = ns["boost_mvs_network"].Network(preprocess=True).eval()
    g = torch.Generator().manual_seed(17)
    for name, buf in net.named_buffers():
        if name.endswith("running_mean"):
            buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
        elif name.endswith("running_var"):
            buf.copy_(torch.rand(buf.shape, generator=g) * 0.5 + 0.75)
    scene = _mvs_scene(9)
    k_best = [2, 1]
    net.view_selection_outputs = {"synth_0": k_best}
    batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()}
    with torch.no_grad(), _ZeroedEmpty():
        ret = net(batch)
        sel = net.forward_view_selection({k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()})
    out = {f"out_{k}": _np(v) for k, v in ret.items()}
    for k in ("all_src_inps", "all_src_exts", "all_src_ixts", "tar_ext", "tar_ixt", "depth_ranges", "rays_0"):
        out[f"in_{k}"] = _np(scene[k])
    out["k_best"] = np.array(k_best, dtype=np.int64)
    out["view_selection"] = np.array(sel["synth_0"], dtype=np.int64)
    for k in ("src_inps", "src_exts", "src_ixts", "near_far"):
        out[f"after_{k}"] = _np(batch[k])
    for name, t in net.state_dict().items():
        out[f"sd_{name}"] = _np(t)
    _save("mvsnerf_chain.npz", out)


def gen_mvs_chain_planes(D):
    """boost_mvsnerf.Network.forward at D = S = 32 (the shipped setting) / 128 (BASELINE config 3) depth planes: same
    network (seed), scene and selection as mvsnerf_chain.npz, so only the outputs are stored (weights of every 8th ray)."""
    ns = load_reference(MVS_CFG, ["enerf.cas_config.k_best", 2, "enerf.cas_config.num_samples", f"[{D}]"])
    torch.manual_seed(13)
    net = ns["boost_mvs_network"].Network(preprocess=True).eval()
    g = torch.Generator().manual_seed(17)
    for name, buf in net.named_buffers():
        if name.endswith("running_mean"):
            buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
        elif name.endswith("running_var"):
            buf.copy_(torch.rand(buf.shape, generator=g) * 0.5 + 0.75)
    scene = _mvs_scene(9)
    net.view_selection_outputs = {"synth_0": [2, 1]}
    batch = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()}
    with torch.no_grad(), _ZeroedEmpty():
        ret = net(batch)
    out = {"out_rgb_level0": _np(ret["rgb_level0"]), "out_depth_level0": _np(ret["depth_level0"]),
           "out_weights_level0_every8": _np(ret["weights_level0"][:, ::8].contiguous()), "planes": np.array([D], dtype=np.int64)}
    _save(f"mvsnerf_chain_d{D}.npz", out)


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if case == "all":
        for c in ("ops", "chain_eval", "chain_pretrain", "single", "mvs_ops", "mvs_chain", "mvs_chain_d32", "mvs_chain_d128"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), c])
    elif case == "ops":
        gen_ops()
    elif case in ("chain_eval", "chain_pretrain"):
        _chain(case)
    elif case == "single":
        gen_single()
    elif case == "mvs_ops":
        gen_mvs_ops()
    elif case == "mvs_chain":
        gen_mvs_chain()
    elif case in ("mvs_chain_d32", "mvs_chain_d128"):
        gen_mvs_chain_planes(int(case.split("_d")[1]))
    else:
        raise SystemExit(f"unknown case {case}")
