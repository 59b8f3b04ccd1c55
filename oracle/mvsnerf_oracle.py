"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the MVSNeRF / boost-MVSNeRF per-frame
rendering path of BoostMVSNeRFs.  Same rules as oracle/enerf_oracle.py:
only tests/, smoke() and bench.py's CPU legs may import it; every function restates one reference
function operation by operation with `cfg` replaced by explicit arguments.

Parity pin: tests/golden/mvsnerf_*.npz, produced by oracle/gen_golden.py from the UNMODIFIED
reference (with the inplace_abn shim of oracle/shims); tests/test_oracle_golden_mvs.py checks this
module against them bit-exactly.

Deliberate, documented deviation: the reference leaves channels 0:3 of the
24-pixel border of the 41-channel volume UNINITIALISED (torch.empty, mvsnerf/network.py:906); the
oracle (and the kernels) define them as 0.  The golden generator zero-fills the same bytes by
pre-seeding the allocator, see gen_golden.py.
"""
import torch
import torch.nn.functional as F

from oracle import enerf_oracle as E

PAD = 24


def proj_mats(src_exts, src_ixts):
    """reference lib/networks/mvsnerf/network.py:1070-1090 (get_proj_mats).
    src_exts (B,V,4,4), src_ixts (B,V,3,3) -> (B,V,3,4); view 0 is the reference (identity)."""
    out_b = []
    for b in range(src_exts.shape[0]):
        mats = []
        ref_inv = None
        for i in range(src_exts.shape[1]):
            full = torch.eye(4, device=src_exts.device)
            k = src_ixts[b, i].clone()
            k[:2] *= 0.25
            full[:3, :4] = k @ src_exts[b, i][:3, :4]
            if i == 0:
                ref_inv = torch.inverse(full)
                mats.append(torch.eye(4, device=src_exts.device))
            else:
                mats.append(full @ ref_inv)
        out_b.append(torch.stack(mats)[:, :3])
    return torch.stack(out_b).float()


def depth_planes(depth_ranges_triple, D):
    """reference lib/networks/boost_mvsnerf/network.py:184-187.  -> near, far (0-dim), planes (1,D)."""
    t = torch.linspace(0., 1., steps=D, device=depth_ranges_triple.device, dtype=torch.float32)
    near, far = depth_ranges_triple.min() * 0.8, depth_ranges_triple.max() * 1.2
    return near, far, (near * (1. - t) + far * t).unsqueeze(0)


def homography_warp(src, proj, planes, pad, grid=None):
    """reference lib/networks/mvsnerf/utils.py:580-630 (homo_warp).  src (B,C,h,w), proj (B,3,4),
    planes (B,D) -> warped (B,C,D,h+2p,w+2p), grid (B,D,(h+2p)*(w+2p),2).  No clamp on z."""
    B, C, h, w = src.shape
    hp, wp = h + 2 * pad, w + 2 * pad
    if grid is None:
        D = planes.shape[1]
        dv = planes[..., None, None].repeat(1, 1, hp, wp)
        rot, trans = proj[:, :, :3], proj[:, :, 3:]
        xs = torch.linspace(0, wp - 1, wp, dtype=torch.float32, device=src.device)
        ys = torch.linspace(0, hp - 1, hp, dtype=torch.float32, device=src.device)
        gy, gx = torch.meshgrid(ys, xs, indexing='ij')
        pix = torch.stack([gx, gy], 0)[None]
        if pad > 0:
            pix = pix - pad
        pix = pix.reshape(1, 2, hp * wp).expand(B, -1, -1)
        pix = torch.cat((pix, torch.ones_like(pix[:, :1])), 1).repeat(1, 1, D)
        cam = rot @ pix + trans / dv.view(B, 1, D * hp * wp)
        g = cam[:, :2] / cam[:, 2:]
        g[:, 0] = g[:, 0] / ((w - 1) / 2) - 1
        g[:, 1] = g[:, 1] / ((h - 1) / 2) - 1
        grid = g.permute(0, 2, 1).view(B, D, hp * wp, 2)
    D = grid.shape[1]
    out = F.grid_sample(src, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    return out.view(B, -1, D, hp, wp), grid


def cost_volume_var_img(imgs, feats, pmats, planes, pad=PAD):
    """reference lib/networks/mvsnerf/network.py:887-942 (build_volume_costvar_img, eval branch).
    imgs (B,V,3,H,W) raw [-1,1]; feats (B,V,32,h,w); pmats (B,V,3,4); planes (B,D)
    -> (B,41,D,h+2p,w+2p): [ref rgb | src rgb x(V-1) | feature variance].  Border of channels 0:3 = 0."""
    B, V, C, h, w = feats.shape
    D = planes.shape[1]
    hp, wp = h + 2 * pad, w + 2 * pad
    ref = F.pad(feats[:, 0], (pad, pad, pad, pad), "constant", 0) if pad > 0 else feats[:, 0]
    vol = torch.zeros((B, 9 + 32, D, hp, wp), device=feats.device, dtype=torch.float)
    small = F.interpolate(imgs.view(B * V, *imgs.shape[2:]), (h, w), mode='bilinear',
                          align_corners=False).view(B, V, -1, h, w).permute(1, 0, 2, 3, 4)
    vol[:, :3, :, pad:h + pad, pad:w + pad] = small[0].unsqueeze(2).expand(-1, -1, D, -1, -1)
    acc = ref.unsqueeze(2).repeat(1, 1, D, 1, 1)
    acc_sq = acc ** 2
    inside = torch.ones((B, V, D, hp, wp), device=feats.device)
    for i in range(1, V):
        warped, grid = homography_warp(feats[:, i], pmats[:, i], planes, pad)
        vol[:, i * 3:(i + 1) * 3], _ = homography_warp(small[i], pmats[:, i], planes, pad, grid=grid)
        g = grid.view(B, 1, D, hp, wp, 2)
        m = (g > -1.0) * (g < 1.0)
        inside[:, i:i + 1] = (m[..., 0] * m[..., 1]).float()
        acc += warped
        acc_sq += warped.pow_(2)
    cnt = 1.0 / torch.sum(inside, dim=1, keepdim=True)
    vol[:, -32:] = acc_sq * cnt - (acc * cnt) ** 2
    return vol


def ray_marcher(rays, S):
    """reference lib/networks/mvsnerf/network.py:945-958.  near/far are ray columns 6,7.
    rays (B,R,8) -> xyz (B,R,S,3), z (B,R,S)."""
    near, far = rays[..., 6:7], rays[..., 7:8]
    t = torch.linspace(0., 1., steps=S, device=rays.device)
    z = near * (1. - t) + far * t
    return rays[..., :3].unsqueeze(2) + rays[..., 3:6].unsqueeze(2) * z.unsqueeze(3), z


def ndc_coordinate(w2c, K, pts, inv_scale, near=2, far=6, pad=0):
    """reference lib/networks/mvsnerf/utils.py:112-146 (get_ndc_coordinate).  pts (R,S,3) -> (R,S,3)."""
    R, S = pts.shape[:2]
    p = torch.matmul(pts.reshape(-1, 3), w2c[:3, :3].t()) + w2c[:3, 3:].reshape(1, 3)
    q = p @ K.t()
    q[:, :2] = (q[:, :2] / q[:, -1:] + 0.0) / inv_scale.reshape(1, 2)
    q[:, 2] = (q[:, 2] - near) / (far - near)
    if pad > 0:
        Wf, Hf = (inv_scale + 1) / 4.0
        q[:, 1] = q[:, 1] * Hf / (Hf + pad * 2) + pad / (Hf + pad * 2)
        q[:, 0] = q[:, 0] * Wf / (Wf + pad * 2) + pad / (Wf + pad * 2)
    return q.view(R, S, 3)


def index_point_feature(volume, ndc):
    """reference lib/networks/mvsnerf/utils.py:357-383 (chunk=-1).  volume (1,8,D,hp,wp), ndc (1,R,S,3) -> (R,S,8)."""
    R, S = ndc.shape[-3:-1]
    g = ndc.view(-1, 1, R, S, 3) * 2 - 1.0
    return F.grid_sample(volume, g, align_corners=True, mode='bilinear')[:, :, 0].permute(2, 3, 0, 1).squeeze()


def color_volume(pts, w2cs, Ks, imgs):
    """reference lib/networks/mvsnerf/utils.py:300-332 (build_color_volume, with_mask=True, no img_feat).
    pts (R,S,3); imgs (1,V,3,H,W) in [0,1] -> (R,S,4V) = per view [rgb, strict in-mask]."""
    _, V, C, H, W = imgs.shape
    inv_scale = torch.tensor([W - 1, H - 1]).to(imgs.device)
    out = torch.empty((*pts.shape[:2], V * 4), device=imgs.device, dtype=torch.float)
    for v in range(V):
        q = ndc_coordinate(w2cs[v], Ks[v].clone(), pts, inv_scale)[None]
        g = q[..., :2] * 2.0 - 1.0
        data = F.grid_sample(imgs[:, v], g, align_corners=True, mode='bilinear', padding_mode='border')
        m = (g > -1.0) * (g < 1.0)
        m = (m[..., 0] * m[..., 1]).float()
        data = torch.cat((data, m.unsqueeze(1)), dim=1)
        out[..., v * 4:v * 4 + 4] = data[0].permute(1, 2, 0)
    return out


def positional_encoding(x, n_freq=10):
    """reference lib/networks/mvsnerf/network.py:24-58 (Embedder.embed): [x, sin(2^k x), cos(2^k x)], frequency-major."""
    freq = (2. ** torch.linspace(0., n_freq - 1, steps=n_freq)).reshape(1, -1, 1).to(x.device)
    rep = x.dim() - 1
    xs = (x.unsqueeze(-2) * freq.view(*[1] * rep, -1, 1)).reshape(*x.shape[:-1], -1)
    return torch.cat((x, torch.sin(xs), torch.cos(xs)), dim=-1)


def mlp_input(pts, ndc, rays_dir, volume, src_inps, src_exts, src_ixts, render_scale=1.0):
    """reference lib/networks/mvsnerf/network.py:961-1001 (rendering + run_network_mvs) with
    renderer.py:111-137 (gen_dir_feature, gen_pts_feats).  pts (R,S,3), ndc (1,R,S,3), rays_dir (1,R,3),
    volume (1,8,D,hp,wp), src_* of the triple (B=1) -> (R,S,86)."""
    norm = torch.norm(rays_dir, dim=-1)
    angle = (rays_dir / norm.unsqueeze(-1)) @ src_exts[0][0][:3, :3].t()               # (1,R,3)
    rgbs = E.unpreprocess(src_inps, render_scale)
    R, S = pts.shape[:2]
    feat = torch.empty((R, S, 20), device=pts.device, dtype=torch.float)
    feat[..., :8] = index_point_feature(volume, ndc)
    feat[..., 8:] = color_volume(pts, src_exts[0], src_ixts[0], rgbs)
    x = torch.cat((positional_encoding(ndc[0]), feat), dim=-1)
    dirs = angle[0][:, None].expand(-1, S, -1)
    return torch.cat([x, dirs], -1)


def render_chain(rays, volume, src_inps, src_exts, src_ixts, near_far, nerf, S, render_scale=1.0):
    """reference lib/networks/boost_mvsnerf/network.py:97-135 (render_rays), B=1."""
    xyz, z = ray_marcher(rays, S)
    B, R = xyz.shape[:2]
    chunk = R // 10
    H0, W0 = src_inps.shape[-2:]
    H, W = int(H0 * render_scale), int(W0 * render_scale)
    inv_scale = torch.tensor([W - 1, H - 1], dtype=torch.float32, device=xyz.device)
    raw = torch.zeros(B, R, S, 4, device=xyz.device)
    mask = torch.zeros(B, R, S, device=xyz.device)
    for i in range(0, R, chunk):
        pts = xyz[:, i:i + chunk]
        ndc = ndc_coordinate(src_exts[0][0], src_ixts[0][0], pts[0], inv_scale, near=near_far.min(),
                             far=near_far.max(), pad=PAD)[None]
        x = mlp_input(pts[0], ndc, rays[:, i:i + chunk, 3:6], volume, src_inps, src_exts, src_ixts, render_scale)
        o = nerf(x)
        raw[:, i:i + chunk] = o.reshape(B, -1, S, o.shape[-1])
        mask[:, i:i + chunk] = E.mask_viewport(pts, src_exts, src_ixts, inv_scale).reshape(B, -1, S)
    return {'net_output': raw, 'z_vals': z, 'mask': mask}


def boost_mvsnerf_forward(net, batch, rc, k_best):
    """reference lib/networks/boost_mvsnerf/network.py:160-211 (Network.forward), B=1 as hard-coded there.
    `net` supplies feature / cost_reg_2 / nerf.  Mutates batch like the reference (near_far, src_*)."""
    inps = batch['all_src_inps']
    B, N = inps.shape[:2]
    I, K = rc.cost_volume_input_views, rc.k_best
    triples = E.view_triples(N, I, inps.device)[k_best]
    D = rc.num_samples[rc.num - 2]
    feats = net.feature(inps)
    bidx = torch.arange(B, device=inps.device).unsqueeze(-1).expand(-1, I)
    per_k = []
    for k in range(K):
        vidx = triples[:, k]
        near, far, planes = depth_planes(batch['depth_ranges'][bidx, vidx], D)
        batch['near_far'] = torch.stack([near, far])
        batch['src_inps'] = inps[bidx, vidx]
        batch['src_exts'] = batch['all_src_exts'][bidx, vidx]
        batch['src_ixts'] = batch['all_src_ixts'][bidx, vidx]
        pm = proj_mats(batch['src_exts'], batch['src_ixts'])
        vol = cost_volume_var_img(batch['src_inps'], feats[bidx, vidx], pm, planes, PAD)
        vol = net.cost_reg_2(vol)
        vol = vol.reshape(1, -1, *vol.shape[2:])
        per_k.append(render_chain(batch['rays_0'], vol, batch['src_inps'], batch['src_exts'], batch['src_ixts'],
                                  batch['near_far'], net.nerf, rc.num_samples[0], rc.render_scale[0]))
    raws = torch.stack([o['net_output'] for o in per_k], dim=1)
    masks = E.merge_masks(torch.stack([o['mask'] for o in per_k], dim=1), K)
    zs = torch.stack([o['z_vals'] for o in per_k], dim=1)
    out = E.composite_blend(raws, masks, zs, rc.white_bkgd)
    return {f'{k}_level0': v for k, v in out.items()}


def visibility_mask_2d(rays, src_exts, src_ixts, H, W, S=128):
    """reference lib/networks/boost_mvsnerf/network.py:23-45 (calc_mask): march S uniform samples,
    volume-render the per-sample visibility score into a 2-D coverage mask (B,R)."""
    xyz, z = ray_marcher(rays, S)
    B = xyz.shape[0]
    inv_scale = torch.tensor([W - 1, H - 1], dtype=torch.float32, device=xyz.device)
    m = E.mask_viewport(xyz, src_exts, src_ixts, inv_scale)
    m = m.reshape(B, -1, S, 1) / S
    m = m.repeat(1, 1, 1, 4)
    return E.composite(m, z, False)['rgb'].mean(-1)


def search_k_best_views(masks, k):
    """reference lib/networks/boost_mvsnerf/network.py:47-71 (= boost_enerf/network.py:71-95): greedy
    coverage search.  masks: list of (B,...) tensors, one per candidate triple -> list of indices."""
    results = []
    prev = torch.ones_like(masks[0])
    HW = masks[0].shape[-2] * masks[0].shape[-1]
    for _ in range(k):
        best_ratio, best = 0, None
        for i in range(len(masks)):
            if i in results:
                continue
            ratio = (masks[i] * prev).sum() / HW
            if ratio > best_ratio:
                best_ratio, best = ratio, i
        if best is None:
            break
        prev = prev * (1 - masks[best])
        results.append(best)
    if results == []:
        results.append(0)
    return results
