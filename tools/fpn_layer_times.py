"""Per-layer timing of the kept FPN (folded, channels-last vs NCHW) on the C2 shape. GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from boostmvsnerfs_b200.modules import FeatureNet
from boostmvsnerfs_b200.inference_plan import folded_copy


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(fmt, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    net = folded_copy(FeatureNet().cuda().eval(), fmt)
    x = torch.randn(6, 3, 544, 960, device="cuda")
    if fmt is not None: x = x.contiguous(memory_format=fmt)
    rows = []
    with torch.no_grad():
        c0a = net.conv0[0](x); rows.append(("conv0.0 3->8 k3 full", t(lambda: net.conv0[0](x))))
        c0 = net.conv0[1](c0a); rows.append(("conv0.1 8->8 k3 full", t(lambda: net.conv0[1](c0a))))
        c1a = net.conv1[0](c0); rows.append(("conv1.0 8->16 k5 s2", t(lambda: net.conv1[0](c0))))
        c1 = net.conv1[1](c1a); rows.append(("conv1.1 16->16 k3 half", t(lambda: net.conv1[1](c1a))))
        c2a = net.conv2[0](c1); rows.append(("conv2.0 16->32 k5 s2", t(lambda: net.conv2[0](c1))))
        c2 = net.conv2[1](c2a); rows.append(("conv2.1 32->32 k3 quarter", t(lambda: net.conv2[1](c2a))))
        q = net.toplayer(c2); rows.append(("toplayer 1x1 32->32", t(lambda: net.toplayer(c2))))
        l1 = net.lat1(c1); rows.append(("lat1 1x1 16->32 half", t(lambda: net.lat1(c1))))
        up = lambda a: F.interpolate(a, scale_factor=2, mode='bilinear', align_corners=True)
        h = up(q) + l1; rows.append(("up(q)+lat1 half", t(lambda: up(q) + l1)))
        l0 = net.lat0(c0); rows.append(("lat0 1x1 8->32 full", t(lambda: net.lat0(c0))))
        f = up(h) + l0; rows.append(("up(h)+lat0 full (32ch)", t(lambda: up(h) + l0)))
        rows.append(("smooth1 32->16 k3 half", t(lambda: net.smooth1(h))))
        rows.append(("smooth0 32->8 k3 full", t(lambda: net.smooth0(f))))
        rows.append(("TOTAL forward", t(lambda: net(x))))
    print(f"--- format={fmt} tf32={tf32}")
    for k, v in rows: print(f"  {k:28s} {v:7.3f} ms")


if __name__ == "__main__":
    run(torch.channels_last, True)
    run(None, True)
