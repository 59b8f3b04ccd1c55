"""2-GPU parity of the sharded single-frame renderer (NCCL) against the single-GPU Network.
Skipped on boxes with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.config import RenderConfig
    from boostmvsnerfs_b200.dist import ShardedFrameRenderer
    from boostmvsnerfs_b200.synth import batch_to, make_scene
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.backends.cudnn.allow_tf32 = False
        rc = RenderConfig.enerf_pretrain(3)          # K=3 on 2 ranks: ragged chain ownership, both levels rendered
        torch.manual_seed(0)
        net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().to(dev)
        net.view_selection_outputs = {"synth_0": [1, 5, 8]}
        scene = make_scene(H=96, W=160, n_views=5, seed=1, smooth=True)
        single = net(batch_to(scene, dev))
        out = ShardedFrameRenderer(net).forward(batch_to(scene, dev))
        worst = {}
        for k in single:
            a, b = out[k].float(), single[k].float()
            assert a.shape == b.shape, (k, a.shape, b.shape)
            worst[k] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_frame_matches_single_gpu():
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        for k, v in ret[r].items():
            # batch-size dependent cuDNN algorithm choices are the only difference
            assert v < 1e-4, f"rank {r} {k}: rel err {v}"
