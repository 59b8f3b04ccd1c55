"""Step-by-step run of the sharded frame under torchrun with flushed progress lines (debugging aid):
    timeout 200 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/dist_debug.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()

    def say(msg):
        print(f"[{time.time() - t0:6.1f}s rank {rank}] {msg}", flush=True)

    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.config import RenderConfig
    from boostmvsnerfs_b200.dist import ShardedFrameRenderer, make_sharded_graph
    from boostmvsnerfs_b200.synth import batch_to, make_scene
    K = int(os.environ.get("DBG_K", "3"))
    H, W = int(os.environ.get("DBG_H", "96")), int(os.environ.get("DBG_W", "160"))
    torch.backends.cudnn.allow_tf32 = os.environ.get("DBG_TF32", "0") == "1"
    rc = RenderConfig.enerf_eval(K)
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().to(dev)
    net.view_selection_outputs = {"synth_0": [1, 5, 8, 3, 2, 7, 0, 9][:K]}
    scene = make_scene(H=H, W=W, n_views=5, seed=1, smooth=True)
    say("single")
    single = {k: v.clone() for k, v in net(batch_to(scene, dev)).items()}
    torch.cuda.synchronize()
    say("warm-up collectives")
    x = torch.ones(4, device=dev)
    dist.all_reduce(x)
    y = torch.empty(4 * world, device=dev)
    dist.all_gather_into_tensor(y, x)
    z = torch.empty(4, device=dev)
    dist.all_to_all_single(z, x[: 4 // world * world].contiguous() if world <= 4 else x)
    torch.cuda.synchronize()
    say("eager sharded")
    out = ShardedFrameRenderer(net).forward(batch_to(scene, dev))
    torch.cuda.synchronize()
    say("eager sharded done: " + ", ".join(f"{k}={float((out[k].float() - single[k].float()).abs().max() / single[k].float().abs().max()):.1e}" for k in single))
    if os.environ.get("DBG_GRAPH", "1") == "1":
        say("graph build")
        fg = make_sharded_graph(net)
        o2 = fg(batch_to(scene, dev))
        torch.cuda.synchronize()
        say("graph done: " + ", ".join(f"{k}={float((o2[k].float() - single[k].float()).abs().max() / single[k].float().abs().max()):.1e}" for k in single))
        for _ in range(3):
            o2 = fg(batch_to(scene, dev))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        e0.record()
        for _ in range(20):
            fg(batch_to(scene, dev), cameras_unchanged=True)
        e1.record()
        torch.cuda.synchronize()
        say(f"graph replay {e0.elapsed_time(e1) / 20:.3f} ms/frame")
    if os.environ.get("DBG_SHARDF", "0") == "1":
        say("sharded features")
        o3 = ShardedFrameRenderer(net, shard_features=True).forward(batch_to(scene, dev))
        torch.cuda.synchronize()
        say("sharded features done: " + ", ".join(f"{k}={float((o3[k].float() - single[k].float()).abs().max() / single[k].float().abs().max()):.1e}" for k in single))
    if os.environ.get("DBG_RELEASE", "1") == "1":
        import gc
        fg = o2 = out = None
        gc.collect()
        torch.cuda.synchronize()
        say("graphs released")
    dist.barrier()
    torch.cuda.synchronize()
    say("destroy")
    dist.destroy_process_group()
    say("bye")


if __name__ == "__main__":
    main()
