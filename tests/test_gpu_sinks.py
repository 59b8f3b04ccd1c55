"""f4 sinks (device PSNR accumulation, uint8 frame conversion) against the numpy restatement of the reference's
Evaluator / Visualizer arithmetic (oracle/sinks_oracle.py)."""
import numpy as np
import pytest
import torch

from boostmvsnerfs_b200.config import RenderConfig
from oracle import sinks_oracle as SO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W", [(64, 96), (544, 960), (33, 47)])
@pytest.mark.parametrize("masked,center", [(False, False), (True, False), (True, True), (False, True)])
def test_frame_psnr_matches_oracle(H, W, masked, center):
    from boostmvsnerfs_b200 import ops
    rs = np.random.RandomState(H + W)
    pred = rs.rand(H, W, 3).astype(np.float32)
    gt = np.clip(pred + 0.05 * rs.randn(H, W, 3), 0, 1).astype(np.float32)
    mask = (rs.rand(H, W) > 0.3).astype(np.uint8) * 2 if masked else None          # the loaders store 0 / >= 1
    want = SO.frame_psnr(pred, gt, mask, eval_center=center)
    m8 = None if mask is None else torch.from_numpy((mask >= 1).astype(np.uint8)).cuda().reshape(-1)
    crop = (int(H * 0.1), int(W * 0.1)) if center else (0, 0)
    sse, cnt = ops.frame_psnr_accumulate(torch.from_numpy(pred).cuda().reshape(-1, 3), torch.from_numpy(gt).cuda().reshape(-1, 3),
                                         H, W, mask=m8, crop=crop)
    got = 10 * np.log10(1.0 / (float(sse.item()) / int(cnt.item())))
    assert abs(got - want) < 1e-9 * max(1.0, abs(want)), (got, want)


def test_frame_to_u8_bit_exact():
    from boostmvsnerfs_b200 import ops
    rs = np.random.RandomState(3)
    rgb = rs.rand(300 * 211, 3).astype(np.float32)
    rgb[:7] = [[0, 1, 0.5]] * 7
    depth = (rs.rand(300 * 211) * 6 + 2).astype(np.float32)
    w_rgb, w_dpt = SO.frame_to_u8(rgb, depth)
    g_rgb, g_dpt, mm = ops.frame_to_u8(torch.from_numpy(rgb).cuda(), torch.from_numpy(depth).cuda())
    assert np.array_equal(g_rgb.cpu().numpy(), w_rgb)
    assert np.array_equal(g_dpt.cpu().numpy(), w_dpt)
    assert float(mm[0]) == float(depth.min()) and float(mm[1]) == float(depth.max())
    only_rgb, none_d, none_m = ops.frame_to_u8(rgb=torch.from_numpy(rgb).cuda())
    assert np.array_equal(only_rgb.cpu().numpy(), w_rgb) and none_d is None and none_m is None


def test_sinks_mirror_evaluator_and_visualizer(tmp_path):
    """PsnrAccumulator / FrameWriter on a rendered frame: same numbers as the restated reference arithmetic."""
    from boostmvsnerfs_b200 import network, sinks
    from boostmvsnerfs_b200.synth import batch_to, make_scene
    rc = RenderConfig.enerf_eval(2)
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().cuda()
    net.view_selection_outputs = {"synth_0": [1, 2]}
    scene = make_scene(H=64, W=96, n_views=4, seed=0, smooth=True)
    batch = batch_to(scene, "cuda")
    out = net(batch)
    rs = np.random.RandomState(0)
    gt = rs.rand(1, 64 * 96, 3).astype(np.float32)
    msk = (rs.rand(1, 64 * 96) > 0.2).astype(np.uint8)
    batch["rgb_1"], batch["msk_1"] = torch.from_numpy(gt), torch.from_numpy(msk)
    ev = sinks.PsnrAccumulator(rc, eval_center=True)
    ev.evaluate(out, batch)
    ev.evaluate(out, batch)
    pred = out["rgb_level1"][0].cpu().numpy().reshape(64, 96, 3)
    want = SO.frame_psnr(pred, gt[0].reshape(64, 96, 3), msk[0].reshape(64, 96), eval_center=True)
    got = ev.summarize()
    assert abs(got["psnr"] - want) < 1e-9 * abs(want) and len(ev.scene_psnrs["synth_level1"]) == 2
    fw = sinks.FrameWriter(rc, result_dir=str(tmp_path))
    rgb, dpt = fw.visualize(out, batch)
    w_rgb, w_dpt = SO.frame_to_u8(pred, out["depth_level1"][0].cpu().numpy().reshape(64, 96))
    assert np.array_equal(rgb, w_rgb) and np.array_equal(dpt, w_dpt)
    assert (tmp_path / "imgs" / "000000_rgb.ppm").stat().st_size == 64 * 96 * 3 + len(b"P6\n96 64\n255\n")
