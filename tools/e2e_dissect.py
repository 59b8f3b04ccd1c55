"""Where the end-to-end frame (pinned host batch in, rgb + depth out) spends what it adds to the device-resident replay:
    A  device batch, replay only            B  A + read_back (D2D staging + D2H on a side stream)
    C  host batch + prefetch, no read-back  D  C + read_back (= bench.py's e2e loop)
C2 workload, CUDA events around 40 frames after 5 warm-ups."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from boostmvsnerfs_b200 import network  # noqa: E402
from boostmvsnerfs_b200.config import RenderConfig  # noqa: E402
from boostmvsnerfs_b200.graph import FrameGraph  # noqa: E402
from boostmvsnerfs_b200.synth import batch_to, make_scene  # noqa: E402


def main():
    wl = WORKLOADS["C2"]
    rc = RenderConfig.enerf_eval(wl["K"])
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().cuda()
    net.view_selection_outputs = {"synth_0": wl["k_best"]}
    net.generate_rays = True
    scene = make_scene(H=wl["H"], W=wl["W"], n_views=wl["n_views"], seed=0)
    host = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in scene.items() if not k.startswith("rays_")} for _ in range(2)]
    dev = batch_to({k: v for k, v in scene.items() if not k.startswith("rays_")}, "cuda")
    fg = FrameGraph(net)
    with torch.no_grad():
        out = fg(dev)
        res = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in ("rgb_level1", "depth_level1")}

        def loop(mode, n):
            if mode in "CD":
                fg.prefetch(host[0])
            for j in range(n):
                if mode in "AB":
                    o = fg(dev, cameras_unchanged=True)
                else:
                    o = fg(host[j % 2])
                    fg.prefetch(host[(j + 1) % 2])
                if mode in "BD":
                    fg.read_back(o, res)
            if mode in "BD":
                fg.wait_read_back()

        for mode in "AaBCDd":
            fg.set_batch_views = mode.isupper()              # lower case: without the three batch['src_*'] gathers
            mode = mode.upper()
            loop(mode, 5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loop(mode, 40)
            e1.record()
            torch.cuda.synchronize()
            print(f"{mode}{'' if fg.set_batch_views else ' (no src_* gathers)'}: {e0.elapsed_time(e1) / 40:.4f} ms per frame")
    fg.close()


if __name__ == "__main__":
    main()
