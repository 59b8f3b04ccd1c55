"""Experiment: K1 (cost volumes) reading fp16 feature maps instead of fp32, C2 shapes.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200 import ops


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


torch.manual_seed(0)
N, K = 6, 4
triples = [[0, 1, 2], [1, 3, 4], [2, 4, 5], [0, 3, 5]]
proj = torch.eye(3, 4, device="cuda").repeat(N, 1, 1)
proj[:, :, 3] = torch.randn(N, 3, device="cuda") * torch.tensor([20.0, 10.0, 0.01], device="cuda")
for name, C, Hs, Ws, D, h, w, shared in (("level 0 (multi)", 32, 136, 240, 64, 68, 120, True), ("level 1", 16, 272, 480, 8, 272, 480, False)):
    f32 = torch.randn(N, C, Hs, Ws, device="cuda").contiguous(memory_format=torch.channels_last)
    f16 = f32.half()
    p = proj.clone(); p[:, 0, 0] = Ws / w; p[:, 1, 1] = Hs / h
    vols = torch.empty((K, D, h, w, C), device="cuda", dtype=torch.float16).permute(0, 4, 1, 2, 3)
    vols2 = torch.empty_like(vols)
    if shared:
        planes = torch.linspace(0.5, 4.0, D, device="cuda")
        run = lambda f, o: ops.cost_volume_var_shared_multi(f, triples, p, planes, h, w, out=o)
    else:
        planes = torch.rand(K, D, h, w, device="cuda") * 3 + 0.5

        def run(f, o):
            for k in range(K):
                ops.cost_volume_var(f, triples[k], p, planes[k], out=o[k])
    run(f32, vols); run(f16, vols2)
    d = (vols.float() - vols2.float()).abs().max().item() / vols.float().abs().max().item()
    print(f"{name}: fp32 maps {t(lambda: run(f32, vols)):7.1f} us   fp16 maps {t(lambda: run(f16, vols2)):7.1f} us   rel diff {d:.2e}")
