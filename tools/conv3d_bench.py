"""Times bmv_conv3d_k3 against cuDNN (TF32 and fp32) at the C2 layer shapes.  Run on a B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200 import ops
from boostmvsnerfs_b200.mlp_pack import pack_conv3d_k3

SHAPES = [("cost_reg_1.conv2", 4, 16, 16, 4, 136, 240, True), ("cost_reg_0.conv2", 4, 16, 16, 32, 34, 60, True),
          ("cost_reg_1.conv0", 4, 16, 8, 8, 272, 480, True), ("cost_reg_1.heads", 4, 8, 9, 8, 272, 480, False),
          ("cost_reg_0.conv0", 4, 32, 8, 64, 68, 120, True), ("cost_reg_0.heads", 4, 8, 9, 64, 68, 120, False)]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for name, N, Cin, Cout, D, H, W, relu in SHAPES:
    x = torch.randn((N, Cin, D, H, W), device="cuda").contiguous(memory_format=torch.channels_last_3d)
    w = torch.randn((Cout, Cin, 3, 3, 3), device="cuda") * 0.1
    wcl = w.contiguous(memory_format=torch.channels_last_3d)
    b = torch.randn(Cout, device="cuda")
    wf = pack_conv3d_k3(w)
    if Cout == 9:       # the merged heads: 8 feature channels + depth logits to two dense tensors
        logits = torch.empty((N, 1, D, H, W), device="cuda")
        t_ours = timeit(lambda: ops.conv3d_k3(x, wf, b, Cout, relu, out2=logits, split=8))
    else:
        t_ours = timeit(lambda: ops.conv3d_k3(x, wf, b, Cout, relu))
    if Cin in (16, 32) and Cout <= 8:
        xh = x.half()
        t_h = timeit(lambda: ops.conv3d_k3(xh, wf, b, Cout, relu, no_tma=True))
        t_t = timeit(lambda: ops.conv3d_k3(xh, wf, b, Cout, relu))
        print(f"{name:18s} fp16 input: register-staged {t_h:8.1f} us   TMA-staged {t_t:8.1f} us")
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timeit(lambda: torch.nn.functional.conv3d(x, wcl, b, padding=1))
    torch.backends.cudnn.allow_tf32 = False
    t_fp32 = timeit(lambda: torch.nn.functional.conv3d(x, wcl, b, padding=1))
    torch.backends.cudnn.allow_tf32 = True
    mb = (x.numel() + N * Cout * D * H * W) * 4 / 1e6
    print(f"{name:18s} ours {t_ours:8.1f} us ({mb / t_ours * 1e3:7.0f} GB/s algorithmic)   cudnn tf32 {t_tf32:8.1f} us   cudnn fp32 {t_fp32:8.1f} us")

for name, N, Cout, D, H, W in [("cost_reg_1.conv1 s2", 4, 16, 8, 272, 480), ("cost_reg_0.conv1 s2", 4, 16, 64, 68, 120)]:
    x = torch.randn((N, 8, D, H, W), device="cuda").contiguous(memory_format=torch.channels_last_3d)
    w = torch.randn((Cout, 8, 3, 3, 3), device="cuda") * 0.1
    wcl = w.contiguous(memory_format=torch.channels_last_3d)
    b = torch.randn(Cout, device="cuda")
    wf = pack_conv3d_k3(w)
    t_ours = timeit(lambda: ops.conv3d_k3(x, wf, b, Cout, True, stride=2))
    t_lib = timeit(lambda: torch.relu_(torch.nn.functional.conv3d(x, wcl, b, stride=2, padding=1)))
    mb = (x.numel() + x.numel() // 8 * 2) * 4 / 1e6
    print(f"{name:24s} ours {t_ours:8.1f} us ({mb / t_ours * 1e3:7.0f} GB/s algorithmic)   cudnn tf32 + relu {t_lib:8.1f} us")

from boostmvsnerfs_b200.mlp_pack import pack_convT3d_k3s2
for name, N, Cin, Cout, D, H, W in [("cost_reg_1.conv11T+add", 4, 16, 8, 4, 136, 240), ("cost_reg_1.conv9T+add", 4, 32, 16, 2, 68, 120),
                                    ("cost_reg_0.conv11T+add", 4, 16, 8, 32, 34, 60), ("cost_reg_0.conv9T+add", 4, 32, 16, 16, 17, 30)]:
    x = torch.randn((N, Cin, D, H, W), device="cuda").contiguous(memory_format=torch.channels_last_3d)
    w = (torch.randn((Cin, Cout, 3, 3, 3), device="cuda") * 0.1)
    wcl = w.contiguous(memory_format=torch.channels_last_3d)
    b = torch.randn(Cout, device="cuda")
    skip = torch.randn((N, Cout, 2 * D, 2 * H, 2 * W), device="cuda").contiguous(memory_format=torch.channels_last_3d)
    wf = pack_convT3d_k3s2(w)
    t_ours = timeit(lambda: ops.convT3d_k3s2_add(x, wf, b, Cout, skip=skip))
    t_lib = timeit(lambda: skip + torch.nn.functional.conv_transpose3d(x, wcl, b, stride=2, padding=1, output_padding=1))
    mb = (x.numel() + 2 * skip.numel()) * 4 / 1e6
    print(f"{name:24s} ours {t_ours:8.1f} us ({mb / t_ours * 1e3:7.0f} GB/s algorithmic)   cudnn tf32 + add {t_lib:8.1f} us")
