"""Times the multi-chain render kernels alone at the C2 shape (K=4 chains, 960x544 rays, S=2, N=6 views): the mma.sync
engine (bmv_render_rays_multi) and the tcgen05 engine (bmv_render_rays_multi_umma), and reports how far their raw outputs
are apart.  The kernels' measurement switches come from the environment (BMV_RM_DEBUG / BMV_RU_DEBUG, read once per
process), so a dissection is a shell loop over processes:

    for f in 0 1 2 4 8; do BMV_RU_DEBUG=$f python tools/render_multi_bench.py --engine umma; done
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boostmvsnerfs_b200 import mlp_pack, ops  # noqa: E402
from boostmvsnerfs_b200.modules import NeRF  # noqa: E402
from boostmvsnerfs_b200.synth import batch_to, make_scene  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="both", choices=("mma", "umma", "both"))
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--H", type=int, default=544)
    ap.add_argument("--W", type=int, default=960)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    H, W, N, K, S, D = a.H, a.W, 6, a.K, 2, 8
    g = torch.Generator(device="cuda").manual_seed(0)
    scene = batch_to(make_scene(H=H, W=W, n_views=N, seed=0), "cuda")
    hv, wv = H // 2, W // 2
    depth = 2.0 + 6.0 * torch.rand(K, hv, wv, device="cuda", generator=g)
    std = 0.3 * torch.rand(K, hv, wv, device="cuda", generator=g)
    nf = torch.stack([torch.full((K, hv, wv), 2.0, device="cuda"), torch.full((K, hv, wv), 8.0, device="cuda")], dim=1).contiguous()
    vols = torch.randn(K, D, hv, wv, 8, device="cuda", generator=g).permute(0, 4, 1, 2, 3)
    feat = torch.randn(N, H, W, 8, device="cuda", generator=g).permute(0, 3, 1, 2)
    rgb4 = torch.zeros(N, H, W, 4, device="cuda")
    rgb4[..., :3] = scene["all_src_inps"][0].permute(0, 2, 3, 1)
    rgb = rgb4.permute(0, 3, 1, 2)[:, :3]
    cams = ops.CameraBlock(scene["all_src_exts"][0], scene["all_src_ixts"][0], scene["tar_ext"][0])
    torch.manual_seed(0)
    nerf = NeRF(feat_ch=11, viewdir_agg=True).cuda().eval()
    packs = {"mma": mlp_pack.pack_nerf_weights_mma(nerf), "umma": mlp_pack.pack_nerf_weights_umma(nerf)}
    table = [(0, 1, 2), (0, 2, 3), (1, 2, 3), (0, 1, 3), (3, 4, 5), (2, 3, 5), (1, 0, 4), (5, 2, 0)]
    triples = [table[k % len(table)] for k in range(K)]
    rays = scene["rays_1"][0]
    outs = {}
    for eng in (("mma", "umma") if a.engine == "both" else (a.engine,)):
        args = (depth, std, nf, rays, H, W, False, S, vols, feat, rgb, cams, triples, packs[eng])
        out = ops.render_rays_multi(*args)
        for _ in range(2):
            ops.render_rays_multi(*args, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            ops.render_rays_multi(*args, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        n = K * rays.shape[0] * S
        flags = os.environ.get("BMV_RU_DEBUG" if eng == "umma" else "BMV_RM_DEBUG", "0")
        print(f"{eng} (debug flags {flags}): {ms * 1e3:.1f} us for {K} chains, {ms * 1e3 / K:.1f} us per chain, "
              f"{n / ms / 1e6:.2f} G samples/s, {ms * 1e-3 * 1.965e9 * 148 / (n / 128):.0f} clk per 128-sample tile per SM")
        outs[eng] = out
    if len(outs) == 2:
        a_, b_ = outs["umma"], outs["mma"]
        err = float((a_["raw"] - b_["raw"]).abs().max()) / float(b_["raw"].abs().max())
        print(f"umma vs mma: raw max err / range {err:.2e}; z equal {torch.equal(a_['z_vals'], b_['z_vals'])}; "
              f"visibility equal {torch.equal(a_['vis_mask'], b_['vis_mask'])}")


if __name__ == "__main__":
    main()
