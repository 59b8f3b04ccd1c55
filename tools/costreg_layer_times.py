"""Per-layer timing of the kept 3-D U-Nets (folded, channels_last_3d) on the C2 shapes. GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200.modules import CostRegNet, MinCostRegNet
from boostmvsnerfs_b200.inference_plan import folded_copy


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, net, x):
    fmt = torch.channels_last_3d
    net = folded_copy(net.cuda().eval(), fmt)
    x = x.cuda().contiguous(memory_format=fmt)
    rows = []
    with torch.no_grad():
        s0 = net.conv0(x); rows.append(("conv0", t(lambda: net.conv0(x)), tuple(s0.shape)))
        a = net.conv1(s0); rows.append(("conv1 s2", t(lambda: net.conv1(s0)), tuple(a.shape)))
        s1 = net.conv2(a); rows.append(("conv2", t(lambda: net.conv2(a)), tuple(s1.shape)))
        b = net.conv3(s1); rows.append(("conv3 s2", t(lambda: net.conv3(s1)), tuple(b.shape)))
        s2 = net.conv4(b); rows.append(("conv4", t(lambda: net.conv4(b)), tuple(s2.shape)))
        y = s2
        if hasattr(net, "conv5"):
            c = net.conv5(s2); rows.append(("conv5 s2", t(lambda: net.conv5(s2)), tuple(c.shape)))
            d = net.conv6(c); rows.append(("conv6", t(lambda: net.conv6(c)), tuple(d.shape)))
            e = net.conv7(d); rows.append(("conv7 T", t(lambda: net.conv7(d)), tuple(e.shape)))
            y = s2 + e; rows.append(("add", t(lambda: s2 + e), ()))
        f = net.conv9(y); rows.append(("conv9 T", t(lambda: net.conv9(y)), tuple(f.shape)))
        y1 = s1 + f; rows.append(("add", t(lambda: s1 + f), ()))
        g = net.conv11(y1); rows.append(("conv11 T", t(lambda: net.conv11(y1)), tuple(g.shape)))
        y0 = s0 + g; rows.append(("add", t(lambda: s0 + g), ()))
        rows.append(("feat_conv 8->8", t(lambda: net.feat_conv(y0)), ()))
        rows.append(("depth_conv 8->1", t(lambda: net.depth_conv(y0)), ()))
        rows.append(("TOTAL", t(lambda: net(x)), ()))
    print("---", name, tuple(x.shape))
    for k, v, sh in rows: print(f"  {k:18s} {v:7.3f} ms  {sh}")


if __name__ == "__main__":
    run("cost_reg_0", MinCostRegNet(32), torch.randn(4, 32, 64, 68, 120))
    run("cost_reg_1", CostRegNet(16), torch.randn(4, 16, 8, 272, 480))
