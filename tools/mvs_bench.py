"""Times the MVSNeRF + boost path (BASELINE config C3 and the shipped D=32 setting) on one GPU.
Informational: C3 is a parity-test case, not the benchmark line (bench.py)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boostmvsnerfs_b200.config import RenderConfig
from boostmvsnerfs_b200.network_mvs import BoostMvsnerfNetwork
from boostmvsnerfs_b200.synth import make_scene, batch_to


class Timer:
    def __init__(self): self.rec = []
    def __call__(self, name): return _S(self, name)
class _S:
    def __init__(self, t, n): self.t, self.n = t, n
    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True); self.e0.record(); return self
    def __exit__(self, *a):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); self.t.rec.append((self.n, self.e0, e1)); return False


def run(D, dtype, H=544, W=960, K=4, steps=3):
    torch.manual_seed(0)
    net = BoostMvsnerfNetwork(preprocess=True, rc=RenderConfig.mvsnerf_eval(K, D)).eval().cuda()
    net.view_selection_outputs = {"synth_0": [0, 7, 12, 19]}
    net.volume_dtype = dtype
    batch = batch_to(make_scene(H=H, W=W, n_views=6, seed=0, render_scales=(1.0,), mvs_near_far_cols=True), "cuda")
    net(dict(batch)); torch.cuda.synchronize()
    tm = Timer(); net.stage_timer = tm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): net(dict(batch))
    e1.record(); torch.cuda.synchronize()
    agg = {}
    for n, a, b in tm.rec: agg[n] = agg.get(n, 0.0) + a.elapsed_time(b) / steps
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"D": D, "volume_dtype": str(dtype), "ms_per_frame": ms, "rays_per_s": H * W / ms * 1e3,
                      "stages_ms": {k: round(v, 2) for k, v in agg.items()},
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    run(32, torch.float32)
    run(128, torch.bfloat16, steps=2)
