"""Per-layer timing of the low-resolution core of the cost regularisers at the C2 shapes: bmv_conv3d_small against the
cuDNN fp16 layer it would replace (K = 4 chains batched)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boostmvsnerfs_b200 import ops  # noqa: E402
from boostmvsnerfs_b200.mlp_pack import pack_conv3d_small  # noqa: E402


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    torch.manual_seed(0)
    layers = [("L1 conv3", 16, 32, 2, (4, 136, 240), False), ("L1 conv4", 32, 32, 1, (2, 68, 120), False),
              ("L1 conv5", 32, 64, 2, (2, 68, 120), False), ("L1 conv6", 64, 64, 1, (1, 34, 60), False),
              ("L1 conv7T", 64, 32, 2, (1, 34, 60), True),
              ("L0 conv3", 16, 32, 2, (32, 34, 60), False), ("L0 conv4", 32, 32, 1, (16, 17, 30), False)]
    for name, cin, cout, stride, dhw, tr in layers:
        x = torch.randn(4, cin, *dhw, device="cuda").half().contiguous(memory_format=torch.channels_last_3d)
        b = torch.randn(cout, device="cuda")
        if tr:
            w = (torch.randn(cin, cout, 3, 3, 3, device="cuda") * 0.05)
            skip = torch.randn(4, cout, *(2 * d for d in dhw), device="cuda").half().contiguous(memory_format=torch.channels_last_3d)
            wf = pack_conv3d_small(w, transposed=True)
            ours = lambda: ops.conv3d_small(x, wf, b, cout, transposed=True, relu=False, skip=skip)
            w16 = w.half().contiguous(memory_format=torch.channels_last_3d)
            lib = lambda: skip + torch.nn.functional.conv_transpose3d(x, w16, b.half(), stride=2, padding=1, output_padding=1)
        else:
            w = (torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05)
            wf = pack_conv3d_small(w)
            ours = lambda: ops.conv3d_small(x, wf, b, cout, stride=stride, relu=True)
            w16 = w.half().contiguous(memory_format=torch.channels_last_3d)
            b16 = b.half()
            lib = lambda: torch.cudnn_convolution_relu(x, w16, b16, (stride,) * 3, (1, 1, 1), (1, 1, 1), 1)
        with torch.no_grad():
            print(f"{name:10s} {cin:2d}->{cout:2d} s{stride} {dhw}: conv3d_small {timeit(ours):6.1f} us   cuDNN fp16 {timeit(lib):6.1f} us")


if __name__ == "__main__":
    main()
