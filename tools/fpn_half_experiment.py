"""Experiment: the FPN layers that stay on cuDNN (conv1.x, conv2.x, toplayer) with fp16 activations / weights.  GPU box only."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200.modules import FeatureNet
from boostmvsnerfs_b200.inference_plan import PlanCache


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


plan = PlanCache().get("fpn", FeatureNet().cuda().eval(), torch.channels_last)
f = plan.fpn
fh = copy.deepcopy(f).half()
z0 = torch.randn(6, 32, 272, 480, device="cuda").contiguous(memory_format=torch.channels_last)
z0h = z0.half()


def mid(m, z):
    c1 = m.conv1[1](m.conv1[0].forward_s2d(z))
    c2 = m.conv2(c1)
    return c1, m.toplayer(c2)


with torch.no_grad():
    c1, q = mid(f, z0)
    c1h, qh = mid(fh, z0h)
    print("rel diff c1 %.2e quarter %.2e" % ((c1h.float() - c1).abs().max().item() / c1.abs().max().item(),
                                              (qh.float() - q).abs().max().item() / q.abs().max().item()))
    print(f"TF32 {t(lambda: mid(f, z0)):7.1f} us   fp16 {t(lambda: mid(fh, z0h)):7.1f} us   "
          f"z0.half() {t(lambda: z0.half()):6.1f} us   c1.float() {t(lambda: c1h.float()):6.1f} us   q.float() {t(lambda: qh.float()):6.1f} us")
