"""Chain-level GPU parity: the drop-in Network.forward against the golden outputs of the
UNMODIFIED reference Network.forward (tests/golden/enerf_chain_*.npz, enerf_single.npz), with the
reference's own state_dict loaded through the reference's parameter names.

cuDNN/cuBLAS stay in the chain (kept modules), so TF32 is switched off for the strict 1e-4 check
(SURVEY.md §7 "Reference nondeterminism"); a second run with torch defaults checks the looser bound.
"""
import json
import os

import numpy as np
import pytest
import torch

from boostmvsnerfs_b200.config import RenderConfig
from conftest import load_golden
from oracle import enerf_oracle as O

pytestmark = pytest.mark.gpu


# Captured-graph replay against the eager call of the same frame on the strict-fp32 route (cuDNN convolutions): normally
# bit-identical, but cuDNN's algorithm choice can differ between the two on some boxes (seen once: 3 of 18432 rgb values
# off by 4e-6 of range).  1e-5 of range still catches what these tests are after — stale buffers, wrong weights, a
# camera or selection mix-up are off by the difference between two frames (>= 1e-4, asserted where it matters).
GRAPH_VS_EAGER = 1e-5


def _report(a, b, what, rtol):
    a = a.detach().float().cpu().numpy()
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = float(np.abs(b).max())
    err = np.abs(a - b)
    bad = err > rtol * scale + rtol * np.abs(b)
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{a.size} outside rtol {rtol}; max abs err {err.max():.3e} "
                           f"scale {scale:.3e}")


def _net_and_batch(g, rc, cls_name):
    from boostmvsnerfs_b200 import network
    if cls_name == "boost":
        net = network.BoostEnerfNetwork(preprocess=True, rc=rc)
        net.view_selection_outputs = {"synth_0": g.np("k_best").tolist()}
    else:
        net = network.EnerfNetwork(rc=rc)
    sd = {k[3:]: g.t(k) for k in g.keys() if k.startswith("sd_")}
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    batch = {k[3:]: g.t(k, "cuda") for k in g.keys() if k.startswith("in_")}
    batch["meta"] = {"scene": ["synth"], "tar_view": torch.tensor([0])}
    return net, batch


@pytest.fixture()
def strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("case,rc", [("chain_eval", RenderConfig.enerf_eval(2)),
                                     ("chain_pretrain", RenderConfig.enerf_pretrain(2))])
def test_boost_forward_vs_reference(case, rc, strict_fp32):
    g = load_golden(f"enerf_{case}.npz")
    net, batch = _net_and_batch(g, rc, "boost")
    out = net(batch)
    expect = sorted(k[4:] for k in g.keys() if k.startswith("out_"))
    assert sorted(out.keys()) == expect
    for k in expect:
        _report(out[k], g.np(f"out_{k}"), f"{case} {k}", 1e-4)
    for k in ("src_inps", "src_exts", "src_ixts"):        # batch mutation contract
        assert np.array_equal(batch[k].cpu().numpy(), g.np(f"after_{k}")), k


def test_single_volume_forward_vs_reference(strict_fp32):
    g = load_golden("enerf_single.npz")
    rc = RenderConfig.enerf_eval(1)
    net, batch = _net_and_batch(g, rc, "single")
    batch["src_inps"], batch["src_exts"], batch["src_ixts"] = (
        batch["all_src_inps"], batch["all_src_exts"], batch["all_src_ixts"])
    out = net(batch)
    for k in [k[4:] for k in g.keys() if k.startswith("out_")]:
        _report(out[k], g.np(f"out_{k}"), f"single {k}", 1e-4)


def test_fused_and_unfused_network_paths_agree(strict_fp32):
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    assert net.fused_mlp and net.mlp_engine == "umma"
    fused = net(dict(batch))
    net.mlp_engine = "mma"
    mma = net(dict(batch))
    net.mlp_engine = "fma"
    fma = net(dict(batch))
    net.fused_mlp = False
    unfused = net(dict(batch))
    for k in fused:
        _report(fused[k], unfused[k].cpu().numpy(), f"tcgen05 fused vs unfused {k}", 2e-5)
        _report(mma[k], unfused[k].cpu().numpy(), f"mma.sync fused vs unfused {k}", 2e-5)
        _report(fma[k], unfused[k].cpu().numpy(), f"fp32-FMA fused vs unfused {k}", 2e-5)


def test_boost_forward_default_tf32_is_close():
    """torch defaults (cuDNN TF32 convs): the kept CNNs add ~1e-3 noise on both sides of any
    comparison; the frame must still agree with the reference to 1e-2."""
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    out = net(batch)
    _report(out["rgb_level1"], g.np("out_rgb_level1"), "rgb (tf32 convs)", 1e-2)


def test_forward_vs_oracle_medium_scene(strict_fp32):
    """A larger seeded scene (N=6, K=4, 128x192): CUDA Network vs the CPU oracle running the SAME
    module weights; also checks sample counts and the visibility-count agreement rate."""
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.synth import make_scene, batch_to
    rc = RenderConfig.enerf_eval(4)
    torch.manual_seed(3)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval()
    net.view_selection_outputs = {"synth_0": [0, 7, 12, 19]}
    scene = make_scene(H=128, W=192, n_views=6, seed=2, smooth=True)
    with torch.no_grad():
        ref = O.boost_enerf_forward(net, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in scene.items()},
                                    rc, torch.tensor([[0, 7, 12, 19]]))
    net = net.cuda()
    out = net(batch_to(scene, "cuda"))
    assert out["rgb_level1"].shape == (1, 128 * 192, 3) and out["weights_level1"].shape == (1, 128 * 192, 2)
    for k in ref:
        _report(out[k], ref[k].numpy(), f"medium {k}", 1e-4)


@pytest.mark.parametrize("case,rc", [("chain_eval", RenderConfig.enerf_eval(2)),
                                     ("chain_pretrain", RenderConfig.enerf_pretrain(2))])
def test_view_selection_matches_reference(case, rc, strict_fp32):
    """forward_view_selection (GPU, batched over all triples, MLP skipped) picks the reference's triples."""
    g = load_golden(f"enerf_{case}.npz")
    net, batch = _net_and_batch(g, rc, "boost")
    assert net.forward_view_selection(batch) == {"synth_0": g.np("view_selection").tolist()}
    assert net.forward_view_selection(batch, max_chains_per_pass=3) == {"synth_0": g.np("view_selection").tolist()}


def test_frame_graph_replay_matches_eager(strict_fp32):
    """CUDA-graph replay (graph.py) == eager forward, across frames with different cameras/images and
    with host-resident batches."""
    from boostmvsnerfs_b200.graph import FrameGraph
    from boostmvsnerfs_b200.synth import make_scene, batch_to
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    fg = FrameGraph(net)
    eager = net(dict(batch))
    replay = {k: v.clone() for k, v in fg(batch).items()}
    # Normally bit-identical.  This test runs the strict-fp32 route, whose convolutions are cuDNN's: on some boxes its
    # algorithm choice differs between the eager call and the captured one (seen: 3 of 18432 rgb values off by 4e-6 of
    # range), so the bar is 1e-5 of range; a stale buffer or a camera mix-up is off by the difference between two frames.
    for k in eager:
        _report(replay[k], eager[k].cpu().numpy(), f"graph vs eager {k}", GRAPH_VS_EAGER)
    _report(replay["rgb_level1"], g.np("out_rgb_level1"), "graph vs reference rgb", 1e-4)
    # a different frame (new images, new cameras) through the SAME captured graph, from host memory
    scene2 = make_scene(H=64, W=96, n_views=4, seed=77, smooth=True, tar_offset=(0.2, -0.05, 0.1))
    eager2 = net(batch_to(scene2, "cuda"))
    replay2 = fg(scene2)
    assert len(fg._cache) == 1
    for k in eager2:
        _report(replay2[k], eager2[k].cpu().numpy(), f"graph vs eager, second frame {k}", GRAPH_VS_EAGER)


def test_graph_prefetch_streams_different_frames(strict_fp32):
    """FrameGraph.prefetch(): the next frame's pinned host batch is uploaded on a copy stream while the current
    frame renders; results equal the un-pipelined calls, frame by frame, with the host running ahead."""
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.graph import FrameGraph
    from boostmvsnerfs_b200.synth import make_scene
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=RenderConfig.enerf_eval(2)).eval().cuda()
    net.view_selection_outputs = {"synth_0": [1, 2]}
    net.generate_rays = True
    pin = lambda sc: {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in sc.items() if not k.startswith("rays_")}
    frames = [pin(make_scene(H=64, W=96, n_views=4, seed=s, smooth=True, tar_offset=(0.02 * s, 0.0, 0.01 * s))) for s in range(4)]
    fg = FrameGraph(net)
    want = [{k: v.clone() for k, v in fg(f).items()} for f in frames]      # un-pipelined (also builds the graph)
    got = []
    fg.prefetch(frames[0])
    for i, f in enumerate(frames):
        out = fg(f)
        if i + 1 < len(frames):
            fg.prefetch(frames[i + 1])
        got.append({k: v.clone() for k, v in out.items()})                  # no host sync inside the loop
    torch.cuda.synchronize()
    # Normally bit-identical; on some boxes differences in the 5th digit were seen between the two loops, so this is
    # a tolerance check.  A frame or camera mix-up is off by the difference between two frames (checked to be large).
    for w, g_ in zip(want, got):
        for k in w:
            _report(g_[k], w[k].cpu().numpy(), f"prefetched frame {k}", 1e-3)
    d01 = (want[0]["rgb_level1"] - want[1]["rgb_level1"]).abs().max().item()
    assert d01 > 1e-2, d01


def test_generated_rays_give_the_same_frame(strict_fp32):
    """net.generate_rays: no rays in the batch, same frame; also through the graph."""
    from boostmvsnerfs_b200.graph import FrameGraph
    g = load_golden("enerf_chain_pretrain.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_pretrain(2), "boost")
    ref = net(dict(batch))
    norays = {k: v for k, v in batch.items() if not k.startswith("rays_")}
    out = net(dict(norays))
    for k in ref:
        _report(out[k], ref[k].cpu().numpy(), f"generated rays {k}", 1e-6)
    net.generate_rays = True
    out2 = FrameGraph(net)(dict(batch))
    for k in ref:
        _report(out2[k], ref[k].cpu().numpy(), f"generated rays, graph {k}", GRAPH_VS_EAGER)


def test_training_mode_and_cpu_are_refused():
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.synth import make_scene
    net = network.BoostEnerfNetwork(preprocess=True, rc=RenderConfig.enerf_eval(2))
    net.view_selection_outputs = {"synth_0": [0, 1]}
    scene = make_scene(H=64, W=96, n_views=4)
    with pytest.raises(RuntimeError):
        net.train()(scene)
    with pytest.raises(RuntimeError):
        net.eval()(scene)          # CPU tensors: there is no CPU path
    with pytest.raises(FileNotFoundError):
        network.BoostEnerfNetwork(preprocess=False, view_selection_file="/nonexistent/view_selection.json")


def test_graph_read_back_pipelines_results(strict_fp32):
    """FrameGraph.read_back(): a frame's results are staged device-to-device and copied to pinned host memory on a
    read-back stream while the next frame renders; every frame's host copy equals that frame's own result."""
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.graph import FrameGraph
    from boostmvsnerfs_b200.synth import make_scene
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=RenderConfig.enerf_eval(2)).eval().cuda()
    net.view_selection_outputs = {"synth_0": [1, 2]}
    net.generate_rays = True
    pin = lambda sc: {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in sc.items() if not k.startswith("rays_")}
    frames = [pin(make_scene(H=64, W=96, n_views=4, seed=s, smooth=True, tar_offset=(0.02 * s, 0.0, 0.01 * s))) for s in range(4)]
    fg = FrameGraph(net)
    want = [{k: v.clone() for k, v in fg(f).items()} for f in frames]
    keys = ("rgb_level1", "depth_level1")
    hosts = [{k: torch.empty(want[0][k].shape, dtype=want[0][k].dtype).pin_memory() for k in keys} for _ in frames]
    fg.prefetch(frames[0])
    for i, f in enumerate(frames):
        out = fg(f)
        if i + 1 < len(frames):
            fg.prefetch(frames[i + 1])
        fg.read_back(out, hosts[i])                                          # no host sync inside the loop
    fg.wait_read_back()
    torch.cuda.synchronize()
    for i in range(len(frames)):
        for k in keys:
            _report(hosts[i][k], want[i][k].cpu().numpy(), f"read-back frame {i} {k}", 1e-3)
    assert (hosts[0]["rgb_level1"] - hosts[1]["rgb_level1"]).abs().max().item() > 1e-2


def test_half_feature_taps_option_matches_default():
    """Network.half_feature_taps (the fused FPN step emits the level-1 maps in fp16, K1 reads them with 8-byte taps) is a
    TF32-class variant of the default path: same frame within 1e-2."""
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    ref = net(dict(batch))
    net.half_feature_taps = True
    net.invalidate_plans() if hasattr(net, "invalidate_plans") else None
    got = net(dict(batch))
    net.half_feature_taps = False
    for k in ("rgb_level1", "depth_level1"):
        _report(got[k], ref[k].detach().cpu().numpy(), f"half feature taps {k}", 1e-2)


def test_fpn_side_stream_overlap_is_bit_identical():
    """Network.overlap_fpn_topdown (the FPN top-down launches on a side stream under the level-0 chain) only changes WHEN
    the level-1 / level-2 maps are produced: eager and captured frames are bit-identical to the single-stream schedule."""
    from boostmvsnerfs_b200.graph import FrameGraph
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    assert net.overlap_fpn_topdown
    for _ in range(2):
        on = net(dict(batch))
    fg = FrameGraph(net)
    on_g = {k: v.clone() for k, v in fg(dict(batch)).items()}
    on_g2 = fg(dict(batch))                                  # replay
    net.overlap_fpn_topdown = False
    off = net(dict(batch))
    for k in off:
        assert torch.equal(on[k], off[k]), f"eager {k}"
        assert torch.equal(on_g[k], off[k]) and torch.equal(on_g2[k], off[k]), f"graph {k}"
    fg.close()


def test_umma_mlp_engine_frame_matches_mma_engine():
    """Network.mlp_engine = 'umma' (the default): all chains rendered by bmv_render_rays_multi_umma (tcgen05); the frame
    agrees with the mma.sync engine to the MLP tolerance (2e-5 of range)."""
    from boostmvsnerfs_b200 import _lib
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    net.mlp_engine = "mma"
    ref = net(dict(batch))
    net.mlp_engine = "umma"
    calls = []
    orig = _lib.call
    _lib.call = lambda name, p, s: (calls.append(name), orig(name, p, s))[1]
    try:
        got = net(dict(batch))
    finally:
        _lib.call = orig
    assert calls.count("bmv_render_rays_multi_umma") == 1 and "bmv_render_rays_multi" not in calls
    for k in ref:
        err = float((got[k] - ref[k]).abs().max()) / float(ref[k].abs().max())
        assert err <= 2e-5, (k, err)


def test_frame_graph_fresh_device_batches_get_their_own_cameras(strict_fp32):
    """`for b in loader: fg(to_cuda(b))`: every frame's camera tensors are fresh device allocations that the caching
    allocator may place at the previous frame's addresses.  The graph must use each frame's own
    cameras — tensor identity is not content identity."""
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.graph import FrameGraph
    from boostmvsnerfs_b200.synth import make_scene, batch_to
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=RenderConfig.enerf_eval(2)).eval().cuda()
    net.view_selection_outputs = {"synth_0": [1, 2]}
    fg = FrameGraph(net)
    ptrs = set()
    for s in range(4):
        scene = make_scene(H=64, W=96, n_views=4, seed=s, smooth=True, tar_offset=(0.05 * s, -0.02 * s, 0.03 * s))
        want = {k: v.clone() for k, v in net(batch_to(scene, "cuda")).items()}
        b = batch_to(scene, "cuda")                      # freed at the end of the iteration -> addresses are re-used
        ptrs.add(b["tar_ext"].data_ptr())
        got = fg(b)
        for k in want:
            _report(got[k], want[k].cpu().numpy(), f"frame {s} {k}", GRAPH_VS_EAGER)
        del b, got
    assert len(fg._cache) == 1


def test_frame_graph_follows_weights_and_flags(strict_fp32):
    """A captured graph bakes in pointers to the folded / packed weight copies and the precision routing: after
    load_state_dict or a routing change the next call must re-capture instead of replaying stale weights
    (ADVICE round 1, medium); the LRU bound keeps the number of live graphs fixed; batch['src_*'] is set like forward."""
    from boostmvsnerfs_b200.graph import FrameGraph
    g = load_golden("enerf_chain_eval.npz")
    net, batch = _net_and_batch(g, RenderConfig.enerf_eval(2), "boost")
    fg = FrameGraph(net, max_entries=2)
    b1 = dict(batch)
    first = {k: v.clone() for k, v in fg(b1).items()}
    for k in ("src_inps", "src_exts", "src_ixts"):
        assert np.array_equal(b1[k].cpu().numpy(), g.np(f"after_{k}")), k
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        changed = {k: (v * 1.25 if k.startswith("nerf_1") and v.dtype.is_floating_point else v) for k, v in sd.items()}
    net.load_state_dict(changed)
    want = net(dict(batch))
    got = fg(dict(batch))
    assert len(fg._cache) == 2
    for k in want:
        _report(got[k], want[k].cpu().numpy(), f"after load_state_dict {k}", GRAPH_VS_EAGER)
    assert (got["rgb_level1"] - first["rgb_level1"]).abs().max().item() > 1e-4
    net.mlp_engine = "fma"
    fg(dict(batch))
    assert len(fg._cache) == 2                            # least recently used graph dropped
    net.load_state_dict(sd)
    net.mlp_engine = "mma"
    back = fg(dict(batch))
    for k in first:
        _report(back[k], first[k].cpu().numpy(), f"weights restored {k}", GRAPH_VS_EAGER)


def test_frame_graph_is_selection_agnostic(strict_fp32):
    """The kernels read the view ids of the K triples from device memory, so ONE captured graph renders frames whose
    view selection differs (a 64-view sequence does not re-capture per view): same results as eager for every selection,
    one cache entry, batch['src_*'] follows the frame's last triple."""
    from boostmvsnerfs_b200 import network
    from boostmvsnerfs_b200.graph import FrameGraph
    from boostmvsnerfs_b200.synth import make_scene, batch_to
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=RenderConfig.enerf_eval(3)).eval().cuda()
    fg = FrameGraph(net)
    scene = make_scene(H=64, W=96, n_views=5, seed=4, smooth=True)
    table = network._combinations(5, 3)
    for sel in ([0, 4, 9], [9, 4, 0], [2, 3, 7], [1, 5, 8]):
        net.view_selection_outputs = {"synth_0": sel}
        want = {k: v.clone() for k, v in net(batch_to(scene, "cuda")).items()}
        b = batch_to(scene, "cuda")
        got = fg(b)
        for k in want:
            _report(got[k], want[k].cpu().numpy(), f"selection {sel} {k}", GRAPH_VS_EAGER)
        assert torch.equal(b["src_exts"], b["all_src_exts"][:, list(table[sel[-1]])])
    assert len(fg._cache) == 1 and next(iter(fg._cache.values()))["agnostic"]
    net.multi_chain_render = False                 # per-chain launches bake their view ids: one graph per selection again
    for sel in ([0, 4, 9], [2, 3, 7]):
        net.view_selection_outputs = {"synth_0": sel}
        want = {k: v.clone() for k, v in net(batch_to(scene, "cuda")).items()}
        got = fg(batch_to(scene, "cuda"))
        for k in want:
            _report(got[k], want[k].cpu().numpy(), f"baked selection {sel} {k}", GRAPH_VS_EAGER)
    assert len(fg._cache) == 3
