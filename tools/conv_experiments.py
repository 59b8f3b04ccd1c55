"""Times call-side variants of the KEPT cuDNN modules (feature_net, cost_reg_{0,1}) on the C2 shapes:
memory format, BN folding, fused conv+bias+relu, cudnn.benchmark, TF32 on/off.  GPU box only."""
import copy
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
from boostmvsnerfs_b200.modules import FeatureNet, CostRegNet, MinCostRegNet, _CBR


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fold(m):
    """deep copy with every conv+BN pair folded into a biased conv"""
    m = copy.deepcopy(m).eval()

    def fold_pair(conv, bn, transposed=False):
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shape = [1] * conv.weight.dim()
        shape[1 if transposed else 0] = -1
        w = conv.weight * s.view(shape)
        b = bn.bias - bn.running_mean * s
        if conv.bias is not None:
            b = b + conv.bias * s
        conv.weight = nn.Parameter(w)
        conv.bias = nn.Parameter(b)

    for name, mod in list(m.named_modules()):
        if isinstance(mod, _CBR):
            fold_pair(mod.conv, mod.bn)
            mod.bn = nn.Identity()
        elif isinstance(mod, nn.Sequential) and len(mod) == 2 and isinstance(mod[0], nn.ConvTranspose3d):
            fold_pair(mod[0], mod[1], transposed=True)
            mod[1] = nn.Identity()
    return m


def main():
    torch.manual_seed(0)
    dev = "cuda"
    x2 = torch.randn(6, 3, 544, 960, device=dev)
    v0 = torch.randn(4, 32, 64, 68, 120, device=dev)
    v1 = torch.randn(4, 16, 8, 272, 480, device=dev)
    nets = {"feature_net": (FeatureNet().to(dev).eval(), x2, torch.channels_last),
            "cost_reg_0": (MinCostRegNet(32).to(dev).eval(), v0, torch.channels_last_3d),
            "cost_reg_1": (CostRegNet(16).to(dev).eval(), v1, torch.channels_last_3d)}
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        for bench in (False, True):
            torch.backends.cudnn.benchmark = bench
            for name, (net, x, cl) in nets.items():
                with torch.no_grad():
                    ref = net(x)
                    ref = ref[0] if isinstance(ref, tuple) else ref
                    res = {}
                    res["base"] = timeit(lambda: net(x))
                    fnet = fold(net)
                    out = fnet(x); out = out[0] if isinstance(out, tuple) else out
                    err = (out - ref).abs().max().item() / ref.abs().max().item()
                    res["fold"] = timeit(lambda: fnet(x))
                    xcl = x.contiguous(memory_format=cl)
                    ncl = copy.deepcopy(net).to(memory_format=cl)
                    res["cl"] = timeit(lambda: ncl(xcl))
                    fcl = copy.deepcopy(fnet).to(memory_format=cl)
                    res["fold+cl"] = timeit(lambda: fcl(xcl))
                    o2 = fcl(xcl); o2 = o2[0] if isinstance(o2, tuple) else o2
                    err2 = (o2 - ref).abs().max().item() / ref.abs().max().item()
                print(f"tf32={tf32} benchmark={bench} {name:12s} " + " ".join(f"{k}={v:7.3f}ms" for k, v in res.items())
                      + f"  fold_relerr={err:.2e} fold+cl_relerr={err2:.2e}", flush=True)


if __name__ == "__main__":
    main()
