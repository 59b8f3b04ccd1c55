"""Experiment: the low-resolution cuDNN layers of the cost regularisers (conv3..conv7) with fp16 activations / weights
(fp32 accumulation) instead of TF32.  C2 shapes.  GPU box only."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200.modules import CostRegNet, MinCostRegNet
from boostmvsnerfs_b200.inference_plan import PlanCache


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, cls, s1_shape in (("cost_reg_1", CostRegNet, (4, 16, 4, 136, 240)), ("cost_reg_0", MinCostRegNet, (4, 16, 32, 34, 60))):
    net = cls(16 if cls is CostRegNet else 32).cuda().eval()
    plan = PlanCache().get("cr", net, torch.channels_last_3d)
    n = plan.net
    s1 = torch.randn(s1_shape, device="cuda").contiguous(memory_format=torch.channels_last_3d)

    def low(m, x):
        s2 = m.conv4(m.conv3(x))
        if m.depth_levels == 3:
            return s2 + m.conv7(m.conv6(m.conv5(s2)))
        return s2
    nh = copy.deepcopy(n).half()
    s1h = s1.half()
    with torch.no_grad():
        ref = low(n, s1)
        got = low(nh, s1h)
        err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"{name}: TF32 {t(lambda: low(n, s1)):7.1f} us   fp16 {t(lambda: low(nh, s1h)):7.1f} us   rel diff {err:.2e}  "
              f"out stride(1)={got.stride(1)}")
