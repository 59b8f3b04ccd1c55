"""cuDNN conv + separate bias/ReLU kernels (what nn.Conv + F.relu gives for channels-last tensors) against
torch.cudnn_convolution_relu (one fused cuDNN call) at the shapes of the layers that stay on cuDNN. GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

CASES = [("fpn.conv1.0 s2d 3x3 32->16", 2, 6, 32, 16, (272, 480), 1), ("fpn.conv1.1 3x3 16->16", 2, 6, 16, 16, (272, 480), 1),
         ("fpn.conv2.0 s2d 3x3 64->32", 2, 6, 64, 32, (136, 240), 1), ("fpn.conv2.1 3x3 32->32", 2, 6, 32, 32, (136, 240), 1),
         ("cr1.conv3 s2 16->32", 3, 4, 16, 32, (4, 136, 240), 2), ("cr1.conv4 32->32", 3, 4, 32, 32, (2, 68, 120), 1),
         ("cr1.conv5 s2 32->64", 3, 4, 32, 64, (2, 68, 120), 2), ("cr1.conv6 64->64", 3, 4, 64, 64, (1, 34, 60), 1),
         ("cr0.conv3 s2 16->32", 3, 4, 16, 32, (32, 34, 60), 2), ("cr0.conv4 32->32", 3, 4, 32, 32, (16, 17, 30), 1)]


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


for name, dims, N, Cin, Cout, spatial, stride in CASES:
    fmt = torch.channels_last if dims == 2 else torch.channels_last_3d
    x = torch.randn((N, Cin) + spatial, device="cuda").contiguous(memory_format=fmt)
    w = (torch.randn((Cout, Cin) + (3,) * dims, device="cuda") * 0.1).contiguous(memory_format=fmt)
    b = torch.randn(Cout, device="cuda")
    conv = F.conv2d if dims == 2 else F.conv3d
    sep = lambda: torch.relu_(conv(x, w, b, stride=stride, padding=1))
    t_sep = timeit(sep)
    try:
        fused = lambda: torch.cudnn_convolution_relu(x, w, b, (stride,) * dims, (1,) * dims, (1,) * dims, 1)
        y = fused()
        err = (y - sep()).abs().max().item() / sep().abs().max().item()
        t_fused = timeit(fused)
        print(f"{name:28s} conv+add+relu {t_sep:7.1f} us   cudnn_convolution_relu {t_fused:7.1f} us   rel diff {err:.1e}  cl={y.stride(1) == 1}")
    except Exception as e:
        print(f"{name:28s} conv+add+relu {t_sep:7.1f} us   cudnn_convolution_relu FAILED: {type(e).__name__}: {str(e)[:80]}")
