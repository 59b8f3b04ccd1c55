"""Per-step timing of the FPN inference plan (FusedTopDownFPN) on the C2 shape. GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200 import ops
from boostmvsnerfs_b200.modules import FeatureNet
from boostmvsnerfs_b200.inference_plan import PlanCache


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


plan = PlanCache().get("fpn", FeatureNet().cuda().eval(), torch.channels_last)
f = plan.fpn if hasattr(plan, "fpn") else plan
x = torch.randn(6, 3, 544, 960, device="cuda").contiguous(memory_format=torch.channels_last)
rows = []
with torch.no_grad():
    steps = []
    c0a = f.conv0[0](x); steps.append(("conv0.0 3->8 k3 full", lambda: f.conv0[0](x)))
    c0 = f.conv0[1](c0a); steps.append(("conv0.1 8->8 k3 full", lambda: f.conv0[1](c0a)))
    c1a = f.conv1[0](c0); steps.append(("conv1.0 8->16 k5 s2 (s2d)", lambda: f.conv1[0](c0)))
    c1 = f.conv1[1](c1a); steps.append(("conv1.1 16->16 k3 half", lambda: f.conv1[1](c1a)))
    c2a = f.conv2[0](c1); steps.append(("conv2.0 16->32 k5 s2 (s2d)", lambda: f.conv2[0](c1)))
    c2 = f.conv2[1](c2a); steps.append(("conv2.1 32->32 k3 quarter", lambda: f.conv2[1](c2a)))
    q = f.toplayer(c2); steps.append(("toplayer 1x1", lambda: f.toplayer(c2)))
    half = ops.fpn_topdown(q, c1, f.lat1.weight, f.lat1.bias); steps.append(("topdown half (ours)", lambda: ops.fpn_topdown(q, c1, f.lat1.weight, f.lat1.bias)))
    full = ops.fpn_topdown(half, c0, f.lat0.weight, f.lat0.bias); steps.append(("topdown full (ours)", lambda: ops.fpn_topdown(half, c0, f.lat0.weight, f.lat0.bias)))
    steps.append(("smooth1 32->16 k3 half", lambda: f.smooth1(half)))
    steps.append(("smooth0 32->8 k3 full", lambda: f.smooth0(full)))
    from boostmvsnerfs_b200.mlp_pack import pack_conv2d_k3_c32
    w1, w0 = pack_conv2d_k3_c32(f.smooth1.weight), pack_conv2d_k3_c32(f.smooth0.weight)
    from boostmvsnerfs_b200.mlp_pack import pack_conv2d_k3_c8
    wst = pack_conv2d_k3_c8(f.conv0[1].conv.weight)
    steps.append(("fused stem conv0.0+conv0.1", lambda: ops.fpn_stem(x, f.conv0[0].conv.weight, f.conv0[0].conv.bias, wst, f.conv0[1].conv.bias)))
    steps.append(("fused topdown+smooth1 half", lambda: ops.fpn_topdown_smooth(q, c1, f.lat1.weight, f.lat1.bias, w1, f.smooth1.bias, 16, True)))
    steps.append(("fused topdown+smooth0 full", lambda: ops.fpn_topdown_smooth(half, c0, f.lat0.weight, f.lat0.bias, w0, f.smooth0.bias, 8, False)))
    steps.append(("TOTAL plan forward", lambda: plan(x)))
    for name, fn in steps:
        print(f"  {name:30s} {t(fn):7.3f} ms")
