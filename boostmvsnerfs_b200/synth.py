"""Synthetic scenes shaped like the reference's collated `batch` dict.

There is no network for datasets, so every test and benchmark renders random images from
random-but-plausible camera rigs.  The batch keys and layouts are the ones
`boost_enerf.Network.forward` consumes (reference lib/networks/boost_enerf/network.py:172-198):
all_src_inps (B,N,3,H,W) in [-1,1], all_src_exts (B,N,4,4) world->cam (OpenCV axes),
all_src_ixts (B,N,3,3) full-resolution pixels, tar_ext, tar_ixt, near_far (B,2),
depth_ranges (B,N,2), rays_{level} (B,H_l*W_l,8) = [origin(3), direction(3), x, y] and `meta`.
"""
import numpy as np
import torch


def _look_at(eye, at):
    """cam->world for an OpenCV camera (x right, y down, z forward) at `eye` looking at `at`."""
    z = at - eye
    z = z / np.linalg.norm(z)
    x = np.cross(np.array([0.0, 1.0, 0.0]), z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, eye
    return c2w


def full_image_rays(tar_ext, tar_ixt, H, W, scale=1.0):
    """Per-pixel rays of a full target image, row-major.

    Same arithmetic as the reference data loader's test-split branch
    (reference lib/datasets/enerf_utils.py:25-31,62-71): float64 numpy,
    direction = [x, y, 1] @ (K^-1)^T @ R_c2w^T (NOT normalised), then one cast to float32.
    """
    ixt = np.array(tar_ixt, dtype=np.float64)
    if scale != 1.0:
        ixt[:2] *= scale
        H, W = int(H * scale), int(W * scale)
    c2w = np.linalg.inv(np.asarray(tar_ext, dtype=np.float64))
    X, Y = np.meshgrid(np.arange(W), np.arange(H))
    pix = np.concatenate((X[:, :, None], Y[:, :, None], np.ones_like(X[:, :, None])), axis=-1)
    dirs = pix @ (np.linalg.inv(ixt).T @ c2w[:3, :3].T)
    orig = np.broadcast_to(c2w[:3, 3][None, None], (H, W, 3))
    rays = np.concatenate((orig, dirs, X[..., None], Y[..., None]), axis=-1)
    return rays.astype(np.float32).reshape(-1, 8)


def make_scene(H=544, W=960, n_views=6, seed=0, near_far=(2.0, 8.0), render_scales=(0.25, 1.0),
               scene="synth", tar_view=0, tar_offset=(0.05, 0.02, 0.0), smooth=False,
               mvs_near_far_cols=False):
    """Build one synthetic frame (B=1).  Generator spec: SURVEY.md §8(d).

    smooth=True replaces white-noise images by low-frequency ones (used by a few parity tests so
    that the CNN features are not pure noise); benchmarks use the default white noise.
    mvs_near_far_cols=True overwrites ray columns 6,7 with near*0.8, far*1.2, which is what the
    MVSNeRF marcher reads there (reference lib/networks/mvsnerf/network.py:947).
    """
    rs = np.random.RandomState(seed)
    if smooth:
        low = rs.uniform(-1, 1, size=(1, n_views, 3, max(H // 16, 2), max(W // 16, 2))).astype(np.float32)
        imgs = torch.nn.functional.interpolate(
            torch.from_numpy(low).view(n_views, 3, low.shape[-2], low.shape[-1]), size=(H, W),
            mode="bilinear", align_corners=True).view(1, n_views, 3, H, W).numpy().copy()
    else:
        imgs = rs.uniform(-1, 1, size=(1, n_views, 3, H, W)).astype(np.float32)
    K = np.array([[0.9 * W, 0, W / 2.0], [0, 0.9 * W, H / 2.0], [0, 0, 1]], dtype=np.float64)
    at = np.array([0.0, 0.0, 5.0])
    exts = []
    for i in range(n_views):
        th = (i - (n_views - 1) / 2.0) * 0.08
        eye = np.array([4.5 * np.sin(th), 0.2 * rs.randn(), 0.3 * (1 - np.cos(th))])
        exts.append(np.linalg.inv(_look_at(eye, at)))
    tar_ext = np.linalg.inv(_look_at(np.array(tar_offset, dtype=np.float64), at))
    batch = {
        "all_src_inps": torch.from_numpy(imgs),
        "all_src_exts": torch.from_numpy(np.stack(exts)[None].astype(np.float32)),
        "all_src_ixts": torch.from_numpy(np.broadcast_to(K, (1, n_views, 3, 3)).astype(np.float32).copy()),
        "tar_ext": torch.from_numpy(tar_ext[None].astype(np.float32)),
        "tar_ixt": torch.from_numpy(K[None].astype(np.float32)),
        "near_far": torch.tensor([list(near_far)], dtype=torch.float32),
        "depth_ranges": torch.tensor([[list(near_far)] * n_views], dtype=torch.float32),
        "meta": {"scene": [scene], "tar_view": torch.tensor([tar_view]), "frame_id": torch.tensor([0])},
    }
    tar_ext32 = batch["tar_ext"][0].numpy()
    tar_ixt32 = batch["tar_ixt"][0].numpy()
    for lvl, s in enumerate(render_scales):
        rays = full_image_rays(tar_ext32, tar_ixt32, H, W, s)
        if mvs_near_far_cols:
            rays[:, 6] = near_far[0] * 0.8
            rays[:, 7] = near_far[1] * 1.2
        batch[f"rays_{lvl}"] = torch.from_numpy(rays[None])
    return batch


def batch_to(batch, device, non_blocking=False):
    """`to_cuda` of the reference (reference lib/utils/data_utils.py:564, run.py:114-116):
    tensors move, `meta` stays on the host."""
    out = {}
    for k, v in batch.items():
        if k == "meta":
            out[k] = v
        elif torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out
