"""A/B timing of the captured C2 frame (FrameGraph replay, device-resident inputs) under routing flags of the network:

    python tools/frame_ab.py overlap_fpn_topdown=0 overlap_fpn_topdown=1 mlp_engine=umma

Each argument is one variant: comma-separated `attr=value` settings applied on top of the defaults.  Prints ms per frame
(CUDA events around 30 replays after 5 warm-ups) for the default and for every variant.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from boostmvsnerfs_b200 import network  # noqa: E402
from boostmvsnerfs_b200.config import RenderConfig  # noqa: E402
from boostmvsnerfs_b200.graph import FrameGraph  # noqa: E402
from boostmvsnerfs_b200.synth import batch_to, make_scene  # noqa: E402


def parse(v):
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return {"true": True, "false": False}.get(v.lower(), v)


def time_variant(net, batch, settings, reps=30):
    old = {k: getattr(net, k) for k in settings}
    for k, v in settings.items():
        setattr(net, k, v)
    fg = FrameGraph(net)
    try:
        for _ in range(5):
            fg(batch, cameras_unchanged=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fg(batch, cameras_unchanged=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    finally:
        fg.close()
        for k, v in old.items():
            setattr(net, k, v)


def main():
    wl_name = os.environ.get("WORKLOAD", "C2")
    wl = WORKLOADS[wl_name]
    rc = RenderConfig.enerf_eval(wl["K"])
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().cuda()
    net.view_selection_outputs = {"synth_0": wl["k_best"]}
    batch = batch_to(make_scene(H=wl["H"], W=wl["W"], n_views=wl["n_views"], seed=0), "cuda")
    with torch.no_grad():
        print(f"{wl_name} default: {time_variant(net, batch, {}):.3f} ms")
        for arg in sys.argv[1:]:
            settings = {kv.split("=")[0]: parse(kv.split("=")[1]) for kv in arg.split(",")}
            print(f"{wl_name} {arg}: {time_variant(net, batch, settings):.3f} ms")


if __name__ == "__main__":
    main()
