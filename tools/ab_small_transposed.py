import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.chdir(ROOT)
import frame_ab
from bench import WORKLOADS
from boostmvsnerfs_b200 import network, inference_plan
from boostmvsnerfs_b200.config import RenderConfig
from boostmvsnerfs_b200.synth import batch_to, make_scene
wl = WORKLOADS["C2"]
rc = RenderConfig.enerf_eval(wl["K"])
torch.manual_seed(0)
net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().cuda()
net.view_selection_outputs = {"synth_0": wl["k_best"]}
batch = batch_to(make_scene(H=wl["H"], W=wl["W"], n_views=wl["n_views"], seed=0), "cuda")
with torch.no_grad():
    for flag in (False, True, False, True):
        inference_plan.MergedHeadsCostReg.small_transposed = flag
        print("small_transposed", flag, f"{frame_ab.time_variant(net, batch, {}):.4f} ms")
