"""Times bmv_mvs_render_umma alone on a synthetic scene (default 256x384, S=128: 12.6 M samples) and prints one line.
The kernel's measurement switches come from the environment (BMV_MR_DEBUG_PANEL, read once per process by the launcher),
so a dissection is a shell loop over processes:

    for f in 0 1048576 2097152 4194304 7340032; do BMV_MR_DEBUG_PANEL=$f python tools/mvs_render_dissect.py; done
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boostmvsnerfs_b200 import mlp_pack, ops  # noqa: E402
from boostmvsnerfs_b200.modules_mvs import MvsNerfMlp  # noqa: E402
from boostmvsnerfs_b200.synth import batch_to, make_scene  # noqa: E402


def main():
    H, W, N, D, S = 256, 384, 4, 128, 128
    scene = batch_to(make_scene(H=H, W=W, n_views=N, seed=5, smooth=True, render_scales=(1.0,), mvs_near_far_cols=True), "cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    vol = torch.randn(D, H // 4 + 48, W // 4 + 48, 8, device="cuda", generator=g).permute(3, 0, 1, 2)
    torch.manual_seed(2)
    packed = mlp_pack.pack_mvs_weights_umma(MvsNerfMlp().cuda().eval())
    rgb4 = torch.cat([scene["all_src_inps"][0], torch.zeros_like(scene["all_src_inps"][0][:, :1])], 1).permute(0, 2, 3, 1).contiguous()
    rays = scene["rays_0"][0]
    args = (rays, S, (1, 0, 3), scene["all_src_exts"][0], scene["all_src_ixts"][0], H, W, 1.6, 9.6, vol, rgb4)
    out = ops.mvs_render(*args, packed)
    for _ in range(2):
        ops.mvs_render(*args, packed, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        ops.mvs_render(*args, packed, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = rays.shape[0] * S
    clk_per_pair = ms * 1e-3 * 1.965e9 / (n / 256 / 148)
    print(f"BMV_MR_DEBUG_PANEL={os.environ.get('BMV_MR_DEBUG_PANEL', '0')}: {ms:.3f} ms for {n / 1e6:.1f} M samples, "
          f"{n * 251.4e3 / ms / 1e9:.0f} TFLOP/s, ~{clk_per_pair:.0f} clk per tile pair")


if __name__ == "__main__":
    main()
