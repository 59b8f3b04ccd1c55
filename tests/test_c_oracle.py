"""The plain-C oracle (oracle/hotpath_oracle.c) against the reference's golden vectors and the torch
oracle: pins the exact fp32 operation order of the visibility test independently of PyTorch."""
import ctypes

import numpy as np
import torch

from oracle import build_c
from oracle import enerf_oracle as O


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def test_c_visibility_matches_reference_bit_exact(ops_golden):
    g = ops_golden
    lib = build_c.load()
    xyz = np.ascontiguousarray(g.np("in_xyz_wide")[0].reshape(-1, 3))
    exts = np.ascontiguousarray(g.np("in_src_exts")[0])
    ixts = np.ascontiguousarray(g.np("in_src_ixts")[0])
    cnt = np.zeros(xyz.shape[0], dtype=np.int32)
    lib.oracle_visibility_count(_fp(xyz), xyz.shape[0], _fp(exts), _fp(ixts), 3, 95.0, 63.0,
                                cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    ref = g.np("mask_wide")[0, :, 0]
    assert np.array_equal(cnt, np.rint(ref * 3).astype(np.int32))
    assert np.array_equal((cnt.astype(np.float32) / np.float32(3)), ref)


def test_c_visibility_matches_torch_oracle_on_boundary_heavy_cloud():
    """points sprayed exactly around the frustum faces of a 960x544 rig"""
    from boostmvsnerfs_b200.synth import make_scene
    lib = build_c.load()
    sc = make_scene(H=544, W=960, n_views=6, seed=4, render_scales=())
    gen = torch.Generator().manual_seed(9)
    pts = (torch.rand(1, 200000, 1, 3, generator=gen) - 0.5) * torch.tensor([16.0, 10.0, 24.0]) + torch.tensor([0, 0, 4.0])
    views = [0, 2, 5]
    exts, ixts = sc["all_src_exts"][:, views].contiguous(), sc["all_src_ixts"][:, views].contiguous()
    ref = O.visibility_count(pts, exts, ixts, torch.tensor([[959.0, 543.0]]))[0].numpy()
    xyz = np.ascontiguousarray(pts.numpy().reshape(-1, 3))
    cnt = np.zeros(xyz.shape[0], dtype=np.int32)
    lib.oracle_visibility_count(_fp(xyz), xyz.shape[0], _fp(exts[0].numpy()), _fp(ixts[0].numpy()), 3, 959.0, 543.0,
                                cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert np.array_equal(cnt, ref)
    assert len(np.unique(cnt)) == 4


def test_c_blend_matches_reference(ops_golden):
    g = ops_golden
    lib = build_c.load()
    raws = np.ascontiguousarray(g.np("in_blend_raws")[0])
    masks = np.ascontiguousarray(g.np("in_blend_masks")[0])
    zs = np.ascontiguousarray(g.np("in_blend_z")[0])
    K, R, S = masks.shape
    rgb, depth, w = np.zeros((R, 3), np.float32), np.zeros(R, np.float32), np.zeros((R, S), np.float32)
    lib.oracle_composite_blend(_fp(raws), _fp(masks), _fp(zs), K, R, S, _fp(rgb), _fp(depth), _fp(w))
    np.testing.assert_allclose(rgb, g.np("blend_rgb")[0], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(depth, g.np("blend_depth")[0], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(w, g.np("blend_weights")[0], rtol=2e-6, atol=2e-7)
