"""Host wall-clock and device time of each step of the cost-regulariser inference plan (C2 shapes). GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200 import ops
from boostmvsnerfs_b200.modules import CostRegNet, MinCostRegNet
from boostmvsnerfs_b200.inference_plan import PlanCache


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(n): fn()
    e1.record()
    w1 = time.perf_counter()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (w1 - w0) / n * 1e3


for name, cls, shape in (("cost_reg_1 (CostRegNet)", CostRegNet, (4, 16, 8, 272, 480)), ("cost_reg_0 (MinCostRegNet)", MinCostRegNet, (4, 32, 64, 68, 120))):
    net = cls(shape[1]).cuda().eval()
    plan = PlanCache().get("cr", net, torch.channels_last_3d)
    x = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    n = plan.net
    with torch.no_grad():
        pk = plan._packed_weights(x.device)
        steps = []
        s0 = ops.conv3d_k3(x, *pk['conv0'], 8, relu=True); steps.append(("conv0 (ours)", lambda: ops.conv3d_k3(x, *pk['conv0'], 8, relu=True)))
        a = ops.conv3d_k3(s0, *pk['conv1'], 16, relu=True, stride=2); steps.append(("conv1 s2 (ours)", lambda: ops.conv3d_k3(s0, *pk['conv1'], 16, relu=True, stride=2)))
        s1 = ops.conv3d_k3(a, *pk['conv2'], 16, relu=True); steps.append(("conv2 (ours)", lambda: ops.conv3d_k3(a, *pk['conv2'], 16, relu=True)))
        b = n.conv3(s1); steps.append(("conv3 s2 (cudnn)", lambda: n.conv3(s1)))
        s2 = n.conv4(b); steps.append(("conv4 (cudnn)", lambda: n.conv4(b)))
        y = s2
        if n.depth_levels == 3:
            steps.append(("conv5-7 + add (cudnn)", lambda: s2 + n.conv7(n.conv6(n.conv5(s2)))))
            y = s2 + n.conv7(n.conv6(n.conv5(s2)))
        y9 = ops.convT3d_k3s2_add(y, *pk['conv9'], 16, skip=s1); steps.append(("conv9T+add (ours)", lambda: ops.convT3d_k3s2_add(y, *pk['conv9'], 16, skip=s1)))
        y11 = ops.convT3d_k3s2_add(y9, *pk['conv11'], 8, skip=s0); steps.append(("conv11T+add (ours)", lambda: ops.convT3d_k3s2_add(y9, *pk['conv11'], 8, skip=s0)))
        logits = torch.empty((shape[0], 1) + tuple(shape[2:]), device="cuda")
        steps.append(("heads (ours)", lambda: ops.conv3d_k3(y11, pk['heads'], None, 9, relu=False, out2=logits, split=8)))
        steps.append(("TOTAL plan", lambda: plan(x)))
        print(name)
        for nm, fn in steps:
            dev, host = t(fn)
            print(f"  {nm:24s} device {dev:7.3f} ms   host {host:7.3f} ms")
