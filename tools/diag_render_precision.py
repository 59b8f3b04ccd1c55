"""Where does the render stage lose accuracy at C2 (white-noise images)?  Runs on the GPU box.

Feeds ONE chain's state (our own depth / std / near_far / regularised volume) to (a) the reference op sequence on the
GPU (oracle functions, IEEE arithmetic) and (b) our kernels — stand-alone K3, K3+MLP engines — and prints the error of
every intermediate: z, visibility, trilinear volume features, bilinear image features (feature / rgb / direction
columns), MLP output.  Usage: python tools/diag_render_precision.py [H W]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from boostmvsnerfs_b200 import network, ops  # noqa: E402
from boostmvsnerfs_b200.config import RenderConfig  # noqa: E402
from boostmvsnerfs_b200.synth import batch_to, make_scene  # noqa: E402
from oracle import enerf_oracle as O  # noqa: E402


def stat(name, a, b, out):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    d = (a - b).abs()
    i = int(d.argmax())
    rng = float(b.abs().max())
    out[name] = {"max_abs": float(d[i]), "range": rng, "rel": float(d[i]) / max(rng, 1e-30), "argmax": i,
                 "mean_abs": float(d.mean()), "n_over_1e-4_of_range": int((d > 1e-4 * rng).sum()), "n": d.numel()}
    print(f"{name:34s} max {float(d[i]):.3e} (rel {float(d[i]) / max(rng, 1e-30):.2e}) mean {float(d.mean()):.2e} "
          f"n>1e-4*range {int((d > 1e-4 * rng).sum())}/{d.numel()} at {i}")


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (544, 960)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    K, kb = 4, [0, 7, 12, 19]
    rc = RenderConfig.enerf_eval(K)
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().cuda()
    net.view_selection_outputs = {"synth_0": kb}
    smooth = len(sys.argv) > 3 and sys.argv[3] == "smooth"
    batch = batch_to(make_scene(H=H, W=W, n_views=6, seed=0, smooth=smooth), "cuda")
    table = network._combinations(6, 3)
    triples = [table[j] for j in kb]
    res = {}
    with torch.no_grad():
        inps = batch["all_src_inps"][0]
        exts, ixts = batch["all_src_exts"][0], batch["all_src_ixts"][0]
        feats = net.forward_feat(inps)
        cams, projs, _ = net._camera_stage(exts, ixts, batch["tar_ext"][0], batch["tar_ixt"][0])
        states = net._chain_levels(feats, projs, batch["near_far"][0], triples, H, W)
        st = states[1]
        rays = batch["rays_1"][0]
        k = 1
        tr = list(triples[k])
        depth, std, nf, vol = st["depth"][k], st["std"][k], st["nf"][k], st["feat_vol"][k]
        # ---- reference op sequence on the GPU, fed with OUR chain state
        rays12 = O.build_rays(depth[None], std[None], nf[None], rays[None], 2.0, False)
        xyz, uvd, z = O.sample_along_depth(rays12, 2, False)
        vf = O.vox_feat(O.normalise_uv(uvd, H, W).reshape(1, -1, 3), vol[None].contiguous())
        f2 = feats["level_2"][tr][None].contiguous()
        feat_rgb = torch.cat((f2, O.unpreprocess(inps[tr][None], 1.0)), dim=2)
        ifeat = O.img_feat(xyz, feat_rgb, exts[tr][None], ixts[tr][None], batch["tar_ext"], 1.0)
        raw_ref = net.nerf_1(vf, ifeat)
        inv_scale = torch.tensor([[W - 1, H - 1]], dtype=torch.float32, device="cuda")
        m_ref = O.mask_viewport(xyz, exts[tr][None], ixts[tr][None], inv_scale).reshape(-1)
        # ---- ours: stand-alone K3
        rgb4 = feats.get("rgb_nhwc4")
        if rgb4 is None:
            rgb4 = inps.new_zeros((inps.shape[0], H, W, 4))
            rgb4[..., :3] = inps.permute(0, 2, 3, 1)
        rgb = rgb4.permute(0, 3, 1, 2)[:, :3]
        o = ops.raygen_sample_fetch(depth, std, nf, rays, H, W, False, 2, vol, feats["level_2"], rgb, cams, tr)
        stat("K3 z_vals", o["z_vals"], z, res)
        stat("K3 vis_mask", o["vis_mask"], m_ref, res)
        stat("K3 vox_feat", o["vox_feat"], vf, res)
        stat("K3 img_feat[feat 0:8]", o["img_feat"][..., :8], ifeat[0][..., :8], res)
        stat("K3 img_feat[rgb 8:11]", o["img_feat"][..., 8:11], ifeat[0][..., 8:11], res)
        stat("K3 img_feat[dir 11:15]", o["img_feat"][..., 11:15], ifeat[0][..., 11:15], res)
        raw_k3 = net.nerf_1(o["vox_feat"][None], o["img_feat"][None])
        stat("torch MLP on K3 feats vs ref", raw_k3, raw_ref, res)
        stat("  .. rgb only", raw_k3[..., :3], raw_ref[..., :3], res)
        stat("  .. sigma only", raw_k3[..., 3], raw_ref[..., 3], res)
        # ---- ours: MLP kernels on the REFERENCE features (isolates the MLP arithmetic)
        raw_fma = ops.nerf_mlp(vf[0].contiguous(), ifeat[0].contiguous(), net._packed_mlp(1, "fma"))
        stat("fp32-FMA MLP kernel on ref feats", raw_fma, raw_ref, res)
        # ---- ours: fused engines
        for eng in ("fma", "mma"):
            r = ops.render_rays(depth, std, nf, rays, H, W, False, 2, vol, feats["level_2"], rgb, cams, tr,
                                net._packed_mlp(1, eng), engine=eng)
            stat(f"fused {eng} raw vs ref", r["raw"], raw_ref, res)
            stat(f"  .. rgb", r["raw"][..., :3], raw_ref.reshape(-1, 2, 4)[..., :3], res)
            stat(f"  .. sigma", r["raw"][..., 3], raw_ref.reshape(-1, 2, 4)[..., 3], res)
            stat(f"fused {eng} raw vs torch-MLP-on-K3", r["raw"], raw_k3, res)
        # how sensitive is the MLP to its inputs?  perturb the reference features by 1e-5 (relative to range)
        g = torch.Generator(device="cuda").manual_seed(0)
        pert = ifeat + 1e-5 * ifeat.abs().max() * torch.randn(ifeat.shape, device="cuda", generator=g)
        stat("ref MLP, img_feat perturbed 1e-5", net.nerf_1(vf, pert), raw_ref, res)
        pertv = vf + 1e-5 * vf.abs().max() * torch.randn(vf.shape, device="cuda", generator=g)
        stat("ref MLP, vox_feat perturbed 1e-5", net.nerf_1(pertv, ifeat), raw_ref, res)
        print("ranges: vf", float(vf.abs().max()), "ifeat feat", float(ifeat[..., :8].abs().max()),
              "raw rgb", float(raw_ref[..., :3].abs().max()), "sigma", float(raw_ref[..., 3].abs().max()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"diag_render_precision_{W}x{H}{'_smooth' if smooth else ''}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
