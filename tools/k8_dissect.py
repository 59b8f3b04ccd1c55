"""Timing of bmv_fpn_topdown_smooth at the C2 shapes (full resolution: 6 x 544 x 960, Cin 8 -> 8; half: 6 x 272 x 480,
Cin 16 -> 16), fp16 lateral input.  Run once per setting of BMV_FF_DEBUG / BMV_FF_MINB (read at the first launch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boostmvsnerfs_b200 import ops  # noqa: E402
from boostmvsnerfs_b200.mlp_pack import pack_conv2d_k3_c32  # noqa: E402


def run(cin, cout, H, W, write_mid, reps=20):
    torch.manual_seed(0)
    prev = torch.randn(6, 32, H // 2, W // 2, device="cuda").contiguous(memory_format=torch.channels_last)
    lat_in = torch.randn(6, cin, H, W, device="cuda").half().contiguous(memory_format=torch.channels_last)
    lat = torch.nn.Conv2d(cin, 32, 1).cuda()
    smooth = torch.nn.Conv2d(32, cout, 3, padding=1).cuda()
    wf = pack_conv2d_k3_c32(smooth.weight)
    flush = torch.empty(64 << 20, device="cuda")
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.fpn_topdown_smooth(prev, lat_in, lat.weight, lat.bias, wf, smooth.bias, cout, write_mid)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


if __name__ == "__main__":
    tag = f"BMV_FF_DEBUG={os.environ.get('BMV_FF_DEBUG', '0')} BMV_FF_MINB={os.environ.get('BMV_FF_MINB', '-')}"
    print(f"{tag}: full {run(8, 8, 544, 960, False):.1f} us, half {run(16, 16, 272, 480, True):.1f} us")
