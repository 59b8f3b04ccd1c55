"""Times bmv_conv3d_k3_umma (TMA + tcgen05 + TMEM) against the TMA-staged mma.sync kernel at the C2 layer shapes
(fp16 inputs, the outputs the inference plan uses).  Run on a B200."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200 import ops
from boostmvsnerfs_b200.mlp_pack import pack_conv3d_k3, pack_conv3d_k3_umma

SHAPES = [  # name, N, Cin, Cout, D, H, W, relu, out fp16
    ("cost_reg_1.heads", 4, 8, 9, 8, 272, 480, False, False), ("cost_reg_0.heads", 4, 8, 9, 64, 68, 120, False, False),
    ("cost_reg_1.conv0", 4, 16, 8, 8, 272, 480, True, True), ("cost_reg_1.conv2", 4, 16, 16, 4, 136, 240, True, False),
    ("cost_reg_0.conv2", 4, 16, 16, 32, 34, 60, True, False)]


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


for name, N, Cin, Cout, D, H, W, relu, oh in SHAPES:
    x = torch.randn((N, Cin, D, H, W), device="cuda").contiguous(memory_format=torch.channels_last_3d).half()
    w = torch.randn((Cout, Cin, 3, 3, 3), device="cuda") * 0.1
    b = torch.randn(Cout, device="cuda")
    wf, wu = pack_conv3d_k3(w), pack_conv3d_k3_umma(w)
    kw = dict(out_dtype=torch.float16) if oh else {}
    if Cout == 9:
        logits = torch.empty((N, 1, D, H, W), device="cuda")
        kw = dict(out2=logits, split=8)
    t_old = timeit(lambda: ops.conv3d_k3(x, wf, b, Cout, relu, **kw))
    t_new = timeit(lambda: ops.conv3d_k3(x, wu, b, Cout, relu, engine="umma", **kw))
    vox = N * D * H * W
    mb = vox * (Cin * 2 + (Cout * 2 if oh else Cout * 4)) / 1e6
    print(f"{name:18s} mma.sync {t_old:7.1f} us   tcgen05 {t_new:7.1f} us   ({mb:6.1f} MB algorithmic: {mb / t_new * 1e3:6.0f} GB/s = "
          f"{mb / t_new * 1e3 / 6550.4:.2f} of the HBM roofline)")
