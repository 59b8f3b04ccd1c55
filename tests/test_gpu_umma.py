"""GPU tests of the tcgen05 (UMMA) render path, csrc/render_umma.cu:
  * bmv_umma_selftest — one 128-row tile through the same descriptor / TMEM / commit / tcgen05.ld helpers,
    against a float64 matmul (isolates the tensor-memory plumbing from the renderer);
  * bmv_render_rays_umma against the fp32-FMA kernel, the mma.sync kernel and fetch + the torch module.
Tolerance: 2e-5 of the output range (split-fp16 operands carry 22 bits; fp32 accumulation)."""
import pytest
import torch

from conftest import load_golden
from test_gpu_parity import H, W, TRIPLE, _cams, close, exact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from boostmvsnerfs_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def g():
    return load_golden("enerf_ops.npz")


@pytest.mark.parametrize("N,K", [(64, 32), (16, 32), (32, 16), (64, 96), (256, 16)])
def test_umma_selftest_matches_float64_matmul(ops, N, K):
    torch.manual_seed(N * 100 + K)
    a = torch.randn(128, K, device="cuda")
    w = torch.randn(N, K)
    d = ops.umma_selftest(a, w)
    ref = (a.double().cpu() @ w.double().T).float()
    close(d, ref, f"tcgen05 tile N={N} K={K} vs float64", rtol=2e-6)
    # rows and columns land where the layout says: a one-hot A row picks one row of W
    a1 = torch.zeros(128, K, device="cuda")
    a1[torch.arange(128), torch.arange(128) % K] = 1.0
    d1 = ops.umma_selftest(a1, w)
    close(d1, w.T[torch.arange(128) % K], "one-hot rows", rtol=2e-6)


@pytest.mark.parametrize("seed", [4, 5])
def test_umma_render_matches_fp32_paths(ops, g, seed):
    from boostmvsnerfs_b200 import mlp_pack
    from boostmvsnerfs_b200.modules import NeRF
    torch.manual_seed(seed)
    net = NeRF(feat_ch=11).eval().cuda()
    for prm in net.parameters():
        if prm.dim() == 1:
            prm.data.normal_(0, 0.2)
    cams = _cams(ops, g)
    args = (g.t("depth_l1", "cuda")[0], g.t("std_l1", "cuda")[0], g.t("near_far_l1", "cuda")[0],
            g.t("in_rays_1", "cuda")[0], H, W, False, 2, g.t("in_regvol1", "cuda")[0], g.t("in_imfeat2", "cuda")[0],
            g.t("in_src_inps", "cuda")[0], cams, TRIPLE)
    o = ops.raygen_sample_fetch(*args, want=("z_vals", "vox_feat", "img_feat", "vis_mask"))
    with torch.no_grad():
        ref = net(o["vox_feat"][None], o["img_feat"][None])[0].view(-1, 2, 4)
    packed = mlp_pack.pack_nerf_weights_umma(net)
    fma = ops.render_rays(*args, mlp_pack.pack_nerf_weights(net))
    mma = ops.render_rays(*args, mlp_pack.pack_nerf_weights_mma(net), engine="mma")
    um = ops.render_rays(*args, packed, engine="umma", want_count=True)
    close(um["raw"], ref, "tcgen05 raw vs fetch + cuBLAS MLP", rtol=2e-5)
    close(um["raw"], fma["raw"], "tcgen05 raw vs fp32-FMA kernel", rtol=2e-5)
    close(um["raw"], mma["raw"], "tcgen05 raw vs mma.sync kernel", rtol=2e-5)
    exact(um["z_vals"], fma["z_vals"], "z_vals")
    exact(um["vis_mask"], fma["vis_mask"], "visibility")
    part = ops.render_rays(*args, packed, engine="umma", ray_begin=1003, n_rays=777)
    exact(part["raw"], um["raw"][1003:1780], "ray sub-range (ragged tiles)")
    one = ops.render_rays(*args, packed, engine="umma", ray_begin=5, n_rays=1)
    exact(one["raw"], um["raw"][5:6], "single ray")
    vol, imf, img = args[8], args[9], args[10]
    vol_cl = vol.permute(1, 2, 3, 0).contiguous().permute(3, 0, 1, 2)
    imf_cl = imf.contiguous(memory_format=torch.channels_last)
    img4 = img.new_zeros((img.shape[0], img.shape[2], img.shape[3], 4))
    img4[..., :3] = img.permute(0, 2, 3, 1)
    args_cl = args[:8] + (vol_cl, imf_cl, img4.permute(0, 3, 1, 2)[:, :3]) + args[11:]
    vec = ops.render_rays(*args_cl, packed, engine="umma")
    close(vec["raw"], um["raw"], "vectorised gather vs scalar gather", rtol=2e-5)
    exact(vec["z_vals"], um["z_vals"], "z_vals (vectorised gather)")
    # repeated launches are deterministic (TMEM / mbarrier phases reset per launch)
    again = ops.render_rays(*args, packed, engine="umma")
    exact(again["raw"], um["raw"], "second launch")


def test_umma_render_rejects_wrong_packing(ops, g):
    from boostmvsnerfs_b200 import mlp_pack
    from boostmvsnerfs_b200._lib import BmvError
    from boostmvsnerfs_b200.modules import NeRF
    net = NeRF(feat_ch=11).eval().cuda()
    cams = _cams(ops, g)
    args = (g.t("depth_l1", "cuda")[0], g.t("std_l1", "cuda")[0], g.t("near_far_l1", "cuda")[0],
            g.t("in_rays_1", "cuda")[0], H, W, False, 2, g.t("in_regvol1", "cuda")[0], g.t("in_imfeat2", "cuda")[0],
            g.t("in_src_inps", "cuda")[0], cams, TRIPLE)
    with pytest.raises(BmvError):
        ops.render_rays(*args, mlp_pack.pack_nerf_weights_mma(net), engine="umma")
