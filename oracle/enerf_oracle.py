"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the ENeRF / boost-ENeRF per-frame
rendering path of BoostMVSNeRFs.

Allowed importers: tests/, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs.  The product package (`boostmvsnerfs_b200/`) never imports this module and fails
loudly when its CUDA library is missing; there is no CPU fallback.

What this is: every function below restates ONE reference function with the reference's own
arithmetic, operation by operation, on torch tensors (the reference's arithmetic dependency is
PyTorch ATen, pinned `pytorch==1.13.1` in reference README.md:25; here torch 2.11), but with the
global `cfg` replaced by explicit arguments.  Each docstring cites the reference lines it follows.

Parity pin: the reference ships no tests or golden vectors for this path, so the pin
is the reference itself, executed in the build container by `oracle/gen_golden.py`; its
inputs/outputs are committed under tests/golden/ and `tests/test_oracle_golden.py` checks this
module against them (bit-exact on CPU).  When /root/reference is present
`tests/test_oracle_vs_reference.py` additionally compares live.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- camera algebra
def proj_mats(src_exts, src_ixts, tar_ext, tar_ixt, src_scale, tar_scale):
    """reference lib/networks/enerf/utils.py:35-55 (get_proj_mats).
    src_exts (B,S,4,4), src_ixts (B,S,3,3), tar_ext (B,4,4), tar_ixt (B,3,3) -> (B,S,3,4)."""
    B, S = src_exts.shape[:2]
    k_src = src_ixts.clone()
    k_src[:, :, :2] *= src_scale
    p_src = k_src @ src_exts[:, :, :3]
    k_tar = tar_ixt.clone()
    k_tar[:, :2] *= tar_scale
    p_tar = k_tar @ tar_ext[:, :3]
    last = torch.zeros((B, 1, 4), dtype=p_tar.dtype, device=p_tar.device)
    last[:, :, 3] = 1
    p_tar_inv = torch.inverse(torch.cat((p_tar, last), dim=1))
    return p_src.view(B, S, 3, 4) @ p_tar_inv.view(B, 1, 4, 4)


# --------------------------------------------------------------------------- depth hypotheses
def depth_planes_first(near_far, D, h, w, depth_inv):
    """reference lib/networks/enerf/utils.py:103-111,149-153 (get_depth_values, depth is None).
    near_far (B,2) -> planes (B,D,h,w), near_far_out (B,2,h,w)."""
    B = near_far.shape[0]
    t = torch.linspace(0., 1., steps=D, device=near_far.device, dtype=torch.float32).view(1, -1).repeat(B, 1)
    if depth_inv:
        disp = 1. / near_far[:, :1] + t * (1. / near_far[:, 1:] - 1. / near_far[:, :1])
        planes = 1. / disp
    else:
        planes = near_far[:, :1] + (near_far[:, 1:] - near_far[:, :1]) * t
    planes = planes.view(B, D, 1, 1).repeat(1, 1, h, w)
    nf = planes[:, [0, -1]]
    if depth_inv:
        nf = 1 / torch.clamp_min(nf, 1e-6)
    return planes.contiguous(), nf


def depth_planes_next(depth, std, near_far, D, up_scale, prev_inv, cur_inv):
    """reference lib/networks/enerf/utils.py:112-153 (get_depth_values, depth given).
    depth/std (B,h0,w0), near_far (B,2,h0,w0) of the previous level -> planes (B,D,h,w),
    near_far_out (B,2,h,w).  Only prev_inv=True exists in the reference (the else branch at
    :130-138 is an ipdb trap)."""
    if not prev_inv:
        raise NotImplementedError("reference traps here (enerf/utils.py:130)")
    if up_scale != 1.:
        size = (int(depth.shape[-2] * up_scale), int(depth.shape[-1] * up_scale))
        depth = F.interpolate(depth[:, None], size=size, align_corners=True, mode='bilinear')[:, 0]
        std = F.interpolate(std[:, None], size=size, align_corners=True, mode='bilinear')[:, 0]
        near_far = F.interpolate(near_far, size=size, align_corners=True, mode='bilinear')
    lo = depth + std           # disparities: larger = nearer
    hi = depth - std
    lo = torch.where(lo > near_far[:, 0], near_far[:, 0], lo)
    hi = torch.where(hi < near_far[:, 1], near_far[:, 1], hi)
    nf = 1. / torch.stack([lo, hi], dim=-1)          # (B,h,w,2) true depth [near, far]
    t = torch.linspace(0., 1., steps=D, device=depth.device, dtype=torch.float32).view(1, 1, 1, -1)
    if cur_inv:
        disp = 1. / nf[..., :1] + t * (1. / nf[..., 1:] - 1. / nf[..., :1])
        planes = (1. / disp).permute(0, 3, 1, 2)
    else:
        planes = (nf[..., :1] + t * (nf[..., 1:] - nf[..., :1])).permute(0, 3, 1, 2)
    nf_out = planes[:, [0, -1]]
    if cur_inv:
        nf_out = 1 / torch.clamp_min(nf_out, 1e-6)
    return planes.contiguous(), nf_out


# --------------------------------------------------------------------------- cost volume
def homography_warp(src_feat, proj, planes):
    """reference lib/networks/enerf/utils.py:57-95 (homo_warp).
    src_feat (B,C,Hs,Ws), proj (B,3,4), planes (B,D,h,w) -> warped (B,C,D,h,w)."""
    B, D, h, w = planes.shape
    C, Hs, Ws = src_feat.shape[1:]
    rot, trans = proj[:, :, :3], proj[:, :, 3:]
    xs = torch.linspace(0, w - 1, w, dtype=torch.float32, device=src_feat.device)
    ys = torch.linspace(0, h - 1, h, dtype=torch.float32, device=src_feat.device)
    gy, gx = torch.meshgrid(ys, xs, indexing='ij')
    pix = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.ones(h * w, dtype=torch.float32,
                                                                 device=src_feat.device)], 0)
    pix = pix[None].expand(B, -1, -1).repeat(1, 1, D)                     # (B,3,D*h*w)
    cam = rot @ pix + trans / planes.view(B, 1, D * h * w)
    uv = cam[:, :2] / torch.clamp_min(cam[:, 2:], 1e-6)
    uv[:, 0] = uv[:, 0] / ((Ws - 1) / 2) - 1
    uv[:, 1] = uv[:, 1] / ((Hs - 1) / 2) - 1
    grid = uv.permute(0, 2, 1).view(B, D, h * w, 2)
    out = F.grid_sample(src_feat, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    return out.view(B, C, D, h, w)


def cost_volume_var(feats, pmats, planes):
    """reference lib/networks/enerf/utils.py:324-351 (build_feature_volume, variance part).
    feats (B,S,C,Hs,Ws), pmats (B,S,3,4), planes (B,D,h,w) -> (B,C,D,h,w)."""
    S = feats.shape[1]
    acc, acc_sq = 0, 0
    for s in range(S):
        warped = homography_warp(feats[:, s], pmats[:, s], planes)
        acc = acc + warped
        acc_sq = acc_sq + warped ** 2
    return acc_sq.div_(S).sub_(acc.div_(S).pow_(2))


def depth_regression(logits, planes, depth_inv):
    """reference lib/networks/enerf/utils.py:722-727 (depth_regression, level >= 0).
    logits, planes (B,D,h,w) -> depth (B,h,w), std (B,h,w)."""
    prob = F.softmax(logits, 1)
    vals = 1. / torch.clamp_min(planes, 1e-6) if depth_inv else planes
    depth = torch.sum(prob * vals, 1)
    var = (prob * (vals - depth.unsqueeze(1)) ** 2).sum(1)
    return depth, torch.clamp_min(var, 1e-10).sqrt()


# --------------------------------------------------------------------------- rays and samples
def build_rays(depth, std, near_far, rays, up_scale, depth_inv):
    """reference lib/networks/enerf/utils.py:392-422 (build_rays).
    depth/std (B,h,w), near_far (B,2,h,w), rays (B,R,8) -> (B,R,12)
    columns [o(3) d(3) x y | ray_near ray_far | vol_near vol_far]."""
    if up_scale != 1.:
        size = (int(depth.shape[-2] * up_scale), int(depth.shape[-1] * up_scale))
        depth = F.interpolate(depth[:, None], size=size, mode='bilinear', align_corners=True)[:, 0]
        std = F.interpolate(std[:, None], size=size, mode='bilinear', align_corners=True)[:, 0]
        near_far = F.interpolate(near_far, size=size, mode='bilinear', align_corners=True)
    if depth_inv:
        a, b = depth + std, depth - std
        a = torch.where(a > near_far[:, 0], near_far[:, 0], a)
        b = torch.where(b < near_far[:, 1], near_far[:, 1], b)
    else:
        a, b = depth - std, depth + std
        a = torch.where(a < near_far[:, 0], near_far[:, 0], a)
        b = torch.where(b > near_far[:, 1], near_far[:, 1], b)
    interval = torch.stack([a, b], dim=-1)                       # (B,H,W,2)
    vol_nf = near_far.permute(0, 2, 3, 1)                        # (B,H,W,2)
    px = rays[:, :, 6:].long()
    B = rays.shape[0]
    interval = torch.stack([interval[i][px[i][:, 1], px[i][:, 0]] for i in range(B)])
    vol_nf = torch.stack([vol_nf[i][px[i][:, 1], px[i][:, 0]] for i in range(B)])
    return torch.cat([rays, interval, vol_nf], dim=-1)


def sample_along_depth(rays12, S, depth_inv):
    """reference lib/networks/enerf/utils.py:424-443 (sample_along_depth).
    -> world_xyz (B,R,S,3), uvd (B,R,S,3) with PIXEL u,v, z_vals (B,R,S)."""
    o, d, px = rays12[..., :3], rays12[..., 3:6], rays12[..., 6:8]
    rn, rf, vn, vf = rays12[..., 8:9], rays12[..., 9:10], rays12[..., 10:11], rays12[..., 11:12]
    if S == 1:
        z = rn + (rf - rn) * 0.5
    else:
        z = rn + (rf - rn) * torch.linspace(0., 1., S, device=rays12.device)[None, None]
    if depth_inv:
        xyz = o[..., None, :] + d[..., None, :] * (1 / torch.clamp_min(z[..., None], 1e-6))
        dn = (vn - z) / torch.clamp_min(vn - vf, 1e-6)
    else:
        xyz = o[..., None, :] + d[..., None, :] * z[..., None]
        dn = (z - vn) / torch.clamp_min(vf - vn, 1e-6)
    uvd = torch.cat([px[..., None, :].repeat(1, 1, S, 1), dn[..., None]], dim=-1)
    return xyz, uvd, z


def normalise_uv(uvd, H, W):
    """reference lib/networks/boost_enerf/network.py:136 (in-place u/(W-1), v/(H-1))."""
    uvd = uvd.clone()
    uvd[..., 0] = uvd[..., 0] / (W - 1)
    uvd[..., 1] = uvd[..., 1] / (H - 1)
    return uvd


def unpreprocess(imgs, render_scale=1.):
    """reference lib/networks/enerf/utils.py:669-676. imgs (B,S,3,H,W) in [-1,1] -> [0,1] resized."""
    img = imgs * 0.5 + 0.5
    B, S, C, H, W = img.shape
    size = (int(H * render_scale), int(W * render_scale))
    img = F.interpolate(img.reshape(B * S, C, H, W), size=size, align_corners=True, mode='bilinear')
    return img.reshape(B, S, C, size[0], size[1])


# --------------------------------------------------------------------------- feature fetch
def vox_feat(uvd_norm, volume):
    """reference lib/networks/enerf/utils.py:458-460 (get_vox_feat).
    uvd_norm (B,P,3) in [0,1], volume (B,C,D,h,w) -> (B,P,C); trilinear, zeros padding."""
    g = uvd_norm[:, None, None] * 2. - 1.
    return F.grid_sample(volume, g, align_corners=True)[:, :, 0, 0].permute(0, 2, 1)


def img_feat(xyz, img_feat_rgb, src_exts, src_ixts, tar_ext, render_scale):
    """reference lib/networks/enerf/utils.py:753-786 (get_img_feat).
    xyz (B,R,S,3), img_feat_rgb (B,V,C,H,W) -> (B,R*S,V,C+4)."""
    B, V, C, H, W = img_feat_rgb.shape
    pts = xyz.reshape(B, -1, 3)
    pts1 = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)
    per_view = []
    for v in range(V):
        cam = (pts1 @ src_exts[:, v].transpose(-1, -2))[..., :3]
        k = src_ixts[:, v].clone()
        k[:, :2] *= render_scale
        pix = cam @ k.transpose(-1, -2)
        g = pix[..., :2] / torch.clamp_min(pix[..., 2:], 1e-6)
        g[..., 0], g[..., 1] = g[..., 0] / (W - 1), g[..., 1] / (H - 1)
        g = g * 2. - 1.
        f = F.grid_sample(img_feat_rgb[:, v], g[:, None], align_corners=True, mode='bilinear',
                          padding_mode='border').permute(0, 2, 3, 1)[:, 0]
        c_tar = tar_ext.inverse()[:, :3, 3]
        c_src = src_exts[:, v].inverse()[:, :3, 3]
        to_t = pts - c_tar[:, None]
        to_s = pts - c_src[:, None]
        to_t = to_t / (torch.norm(to_t, dim=-1, keepdim=True) + 1e-6)
        to_s = to_s / (torch.norm(to_s, dim=-1, keepdim=True) + 1e-6)
        diff = to_t - to_s
        diff_n = torch.norm(diff, dim=-1, keepdim=True)
        dot = torch.sum(to_t * to_s, dim=-1, keepdim=True)
        per_view.append(torch.cat([f, diff / torch.clamp(diff_n, min=1e-6), dot], dim=-1))
    return torch.stack(per_view, -2)


# --------------------------------------------------------------------------- 3-D visibility
def ndc_coords(xyz, src_ext, src_ixt, inv_scale):
    """reference lib/networks/enerf/utils.py:490-508 (get_ndc_coords). xyz (B,R,S,3) -> (B,R,S,3)."""
    B, R, S = xyz.shape[:3]
    rot, trans = src_ext[:, :3, :3], src_ext[:, :3, 3]
    cam = torch.bmm(xyz.reshape(B, -1, 3), rot.transpose(1, 2))
    cam.add_(trans.view(B, 1, 3))
    pix = cam.bmm(src_ixt.transpose(1, 2))
    pix[:, :, :2].div_(pix[:, :, -1:]).div_(inv_scale.view(B, 1, 2))
    return pix.view(B, R, S, 3)


def visibility_count(xyz, src_exts, src_ixts, inv_scale):
    """Integer form of mask_viewport: number of views whose frustum contains each sample.
    -> (B, R*S) int32."""
    B, R, S = xyz.shape[:3]
    cnt = torch.zeros((B, R * S), dtype=torch.int32, device=xyz.device)
    for v in range(src_exts.shape[1]):
        q = ndc_coords(xyz, src_exts[:, v], src_ixts[:, v], inv_scale)
        inside = (q[..., 0] >= 0) & (q[..., 0] <= 1) & (q[..., 1] >= 0) & (q[..., 1] <= 1) & (q[..., 2] > 0)
        cnt += inside.view(B, R * S).int()
    return cnt


def mask_viewport(xyz, src_exts, src_ixts, inv_scale):
    """reference lib/networks/enerf/utils.py:510-520 (mask_viewport) -> (B,R*S,1) fp32 in {0,1/V,..,1}."""
    V = src_exts.shape[1]
    B, R, S = xyz.shape[:3]
    m = torch.zeros((B, R * S, 1), device=xyz.device)
    for v in range(V):
        q = ndc_coords(xyz, src_exts[:, v], src_ixts[:, v], inv_scale)
        inside = (q[..., 0] >= 0) & (q[..., 0] <= 1) & (q[..., 1] >= 0) & (q[..., 1] <= 1) & (q[..., 2] > 0)
        m += inside.view(B, R * S, 1)
    m /= V
    return m


# --------------------------------------------------------------------------- compositing
def composite(raw, z_vals, white_bkgd=False):
    """reference lib/networks/enerf/utils.py:605-637 (raw2outputs). raw (...,S,4), z (...,S)."""
    alpha = 1. - torch.exp(-raw[..., 3])
    T = torch.cumprod(1. - alpha + 1e-10, dim=-1)[..., :-1]
    T = torch.cat([torch.ones_like(alpha[..., 0:1]), T], dim=-1)
    w = alpha * T
    rgb = torch.sum(w[..., None] * raw[..., :3], -2)
    depth = None
    if z_vals is not None:
        w = F.softmax(w, dim=-1)
        depth = torch.sum(w * z_vals, -1)
    if white_bkgd:
        rgb = rgb + (1. - torch.sum(w, -1)[..., None])
    return {'rgb': rgb, 'depth': depth, 'weights': w}


def merge_masks(masks, K):
    """reference lib/networks/boost_enerf/network.py:167-168. masks (B,K,R,S) -> normalised over K."""
    tot = masks.unsqueeze(1).sum(2)
    return torch.where(tot > 0, masks / tot, 1 / K)


def composite_blend(raws, masks, z_vals, white_bkgd=False):
    """reference lib/networks/enerf/utils.py:639-667 (raw2outputs_blend).
    raws (B,K,R,S,4), masks (B,K,R,S) already normalised over K, z_vals (B,K,R,S)."""
    B, K, R, S = raws.shape[:4]
    masks = masks.view(B, K, R, S)
    alpha_k = 1. - torch.exp(-raws[..., 3])
    alpha = torch.sum(alpha_k * masks, dim=1)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[..., 0:1]), 1 - alpha], dim=-1), dim=-1)[..., :-1]
    w = alpha * T
    rgb = torch.sum((T.unsqueeze(1).repeat(1, K, 1, 1) * alpha_k * masks)[..., None] * raws[..., :3], -2)
    rgb = torch.sum(rgb, dim=1)
    depth = None
    if z_vals is not None:
        w = F.softmax(w, dim=-1)
        depth = torch.sum(w * z_vals.mean(1), -1)
    if white_bkgd:
        raise NotImplementedError
    return {'rgb': rgb, 'depth': depth, 'weights': w}


# --------------------------------------------------------------------------- orchestration
def render_chain(rays12, volume, im_feat, src_inps, src_exts, src_ixts, tar_ext, nerf, level, rc):
    """reference lib/networks/boost_enerf/network.py:123-161 (render_rays + batchify_rays_for_mlp)
    for one cost-volume chain.  Returns net_output (B,R,S,4), z_vals (B,R,S), mask (B,R,S)."""
    S = rc.num_samples[level]
    rs = rc.render_scale[level]
    H, W = int(src_inps.shape[-2] * rs), int(src_inps.shape[-1] * rs)
    outs = {'net_output': [], 'z_vals': [], 'mask': []}
    for r0 in range(0, rays12.shape[1], rc.chunk_size):
        chunk = rays12[:, r0:r0 + rc.chunk_size]
        xyz, uvd, z = sample_along_depth(chunk, S, rc.depth_inv[level])
        B = xyz.shape[0]
        rgbs = unpreprocess(src_inps, rs)
        up = rs / rc.im_ibr_scale[level]
        feat = im_feat
        if up != 1.:
            b, s, c, fh, fw = feat.shape
            feat = F.interpolate(feat.reshape(b * s, c, fh, fw), size=(int(fh * up), int(fw * up)),
                                 align_corners=True, mode='bilinear').view(b, s, c, int(fh * up), int(fw * up))
        feat_rgb = torch.cat((feat, rgbs), dim=2)
        vf = vox_feat(normalise_uv(uvd, H, W).reshape(B, -1, 3), volume)
        ifeat = img_feat(xyz, feat_rgb, src_exts, src_ixts, tar_ext, rs)
        net = nerf(vf, ifeat)
        net = net.reshape(B, -1, S, net.shape[-1])
        inv_scale = torch.tensor([W - 1, H - 1], dtype=torch.float32, device=net.device).unsqueeze(0).expand(B, -1)
        m = mask_viewport(xyz, src_exts, src_ixts, inv_scale).reshape(B, -1, S)
        outs['net_output'].append(net)
        outs['z_vals'].append(z)
        outs['mask'].append(m)
    return {k: torch.cat(v, dim=1) for k, v in outs.items()}


def view_triples(n_views, per_volume, device=None):
    """reference lib/networks/boost_enerf/network.py:176 — lexicographic combinations table."""
    return torch.combinations(torch.arange(n_views), per_volume).to(device)


def boost_enerf_forward(net, batch, rc, k_best, internals=None):
    """reference lib/networks/boost_enerf/network.py:172-237 (Network.forward).
    `net` supplies the kept NN modules: forward_feat-compatible `feature_net`, `cost_reg_{i}`,
    `nerf_{i}` (module objects are shared with the product so both sides use identical weights).
    `k_best`: (B,K) long tensor of indices into the triples table.  Mutates `batch` like the
    reference does (src_inps/src_exts/src_ixts of the last triple).
    internals: optional dict that receives, per rendered level, the per-chain visibility scores
    `masks_level{i}` (B,K,R,S) before merge_masks, the sample positions `xyz_level{i}` (B,K,R,S,3) and `triples`
    — what a parity test needs to tell a 1-ulp frustum-edge flip from an arithmetic error."""
    inps = batch['all_src_inps']
    B, N = inps.shape[:2]
    I, K = rc.cost_volume_input_views, rc.k_best
    triples = view_triples(N, I, inps.device)[k_best]                   # (B,K,I)
    x = inps.view(B * N, *inps.shape[2:])
    f2, f1, f0 = net.feature_net(x)
    Hh, Ww = inps.shape[-2:]
    feats = {'level_2': f0.reshape(B, N, f0.shape[1], Hh, Ww),
             'level_1': f1.reshape(B, N, f1.shape[1], Hh // 2, Ww // 2),
             'level_0': f2.reshape(B, N, f2.shape[1], Hh // 4, Ww // 4)}
    depth, std, nf = [None] * K, [None] * K, [None] * K
    bidx = torch.arange(B, device=inps.device).unsqueeze(-1).expand(-1, I)
    ret = {}
    for i in range(rc.num):
        per_k = []
        for k in range(K):
            vidx = triples[:, k]
            batch['src_inps'] = inps[bidx, vidx]
            batch['src_exts'] = batch['all_src_exts'][bidx, vidx]
            batch['src_ixts'] = batch['all_src_ixts'][bidx, vidx]
            D = rc.volume_planes[i]
            vs = rc.volume_scale[i]
            h, w = int(Hh * vs), int(Ww * vs)
            if depth[k] is None:
                planes, nf[k] = depth_planes_first(batch['near_far'], D, h, w, rc.depth_inv[i])
            else:
                planes, nf[k] = depth_planes_next(depth[k], std[k], nf[k], D, vs / rc.volume_scale[i - 1],
                                                  rc.depth_inv[i - 1], rc.depth_inv[i])
            pm = proj_mats(batch['src_exts'], batch['src_ixts'], batch['tar_ext'], batch['tar_ixt'],
                           rc.im_feat_scale[i], vs)
            vol = cost_volume_var(feats[f'level_{i}'][bidx, vidx], pm, planes)
            vol, logits = getattr(net, f'cost_reg_{i}')(vol)
            depth[k], std[k] = depth_regression(logits, planes, rc.depth_inv[i])
            if not rc.render_if[i]:
                continue
            rays12 = build_rays(depth[k], std[k], nf[k], batch[f'rays_{i}'], rc.render_scale[i] / vs,
                                rc.depth_inv[i])
            lvl = rc.render_im_feat_level[i]
            per_k.append(render_chain(rays12, vol, feats[f'level_{lvl}'][bidx, vidx], batch['src_inps'],
                                      batch['src_exts'], batch['src_ixts'], batch['tar_ext'],
                                      getattr(net, f'nerf_{i}'), i, rc))
            if internals is not None:
                per_k[-1]['xyz'] = sample_along_depth(rays12, rc.num_samples[i], rc.depth_inv[i])[0]
        if not rc.render_if[i]:
            continue
        if internals is not None:
            internals[f'masks_level{i}'] = torch.stack([o['mask'] for o in per_k], dim=1)
            internals[f'xyz_level{i}'] = torch.stack([o['xyz'] for o in per_k], dim=1)
            internals['triples'] = triples
        raws = torch.stack([o['net_output'] for o in per_k], dim=1)
        masks = merge_masks(torch.stack([o['mask'] for o in per_k], dim=1), K)
        zs = torch.stack([o['z_vals'] for o in per_k], dim=1)
        out = composite_blend(raws, masks, zs, rc.white_bkgd)
        out['depth_mvs'] = 1. / depth[0] if rc.depth_inv[i] else depth[0]
        out['std'] = std[0]
        ret.update({f'{key}_level{i}': val for key, val in out.items()})
    return ret


def enerf_forward(net, batch, rc):
    """reference lib/networks/enerf/network.py:76-113 (single-volume ENeRF Network.forward)."""
    inps = batch['src_inps']
    B, N = inps.shape[:2]
    x = inps.view(B * N, *inps.shape[2:])
    f2, f1, f0 = net.feature_net(x)
    Hh, Ww = inps.shape[-2:]
    feats = {'level_2': f0.reshape(B, N, f0.shape[1], Hh, Ww),
             'level_1': f1.reshape(B, N, f1.shape[1], Hh // 2, Ww // 2),
             'level_0': f2.reshape(B, N, f2.shape[1], Hh // 4, Ww // 4)}
    depth = std = nf = None
    ret = {}
    for i in range(rc.num):
        D, vs = rc.volume_planes[i], rc.volume_scale[i]
        h, w = int(Hh * vs), int(Ww * vs)
        if depth is None:
            planes, nf = depth_planes_first(batch['near_far'], D, h, w, rc.depth_inv[i])
        else:
            planes, nf = depth_planes_next(depth, std, nf, D, vs / rc.volume_scale[i - 1],
                                           rc.depth_inv[i - 1], rc.depth_inv[i])
        pm = proj_mats(batch['src_exts'], batch['src_ixts'], batch['tar_ext'], batch['tar_ixt'],
                       rc.im_feat_scale[i], vs)
        vol = cost_volume_var(feats[f'level_{i}'], pm, planes)
        vol, logits = getattr(net, f'cost_reg_{i}')(vol)
        depth, std = depth_regression(logits, planes, rc.depth_inv[i])
        if not rc.render_if[i]:
            continue
        rays12 = build_rays(depth, std, nf, batch[f'rays_{i}'], rc.render_scale[i] / vs, rc.depth_inv[i])
        lvl = rc.render_im_feat_level[i]
        o = render_chain(rays12, vol, feats[f'level_{lvl}'], inps, batch['src_exts'], batch['src_ixts'],
                         batch['tar_ext'], getattr(net, f'nerf_{i}'), i, rc)
        out = composite(o['net_output'], o['z_vals'], rc.white_bkgd)
        out['depth_mvs'] = 1. / depth if rc.depth_inv[i] else depth
        out['std'] = std
        ret.update({f'{key}_level{i}': val for key, val in out.items()})
    return ret


# --------------------------------------------------------------------------- view selection (pre-process)
def calc_mask(net, triple, batch, rc):
    """reference lib/networks/boost_enerf/network.py:22-69 (calc_mask): one full single-volume cascade
    for the given triple; the per-sample visibility score is volume-rendered into a 2-D coverage
    mask per rendered level.  (The reference also evaluates the MLP and discards its output.)"""
    inps = batch['all_src_inps'][:, triple]
    exts, ixts = batch['all_src_exts'][:, triple], batch['all_src_ixts'][:, triple]
    B, N = inps.shape[:2]
    Hh, Ww = inps.shape[-2:]
    f2, f1, f0 = net.feature_net(inps.reshape(B * N, *inps.shape[2:]))
    feats = {'level_1': f1.reshape(B, N, f1.shape[1], Hh // 2, Ww // 2),
             'level_0': f2.reshape(B, N, f2.shape[1], Hh // 4, Ww // 4)}
    depth = std = nf = None
    out = {}
    for i in range(rc.num):
        D, vs = rc.volume_planes[i], rc.volume_scale[i]
        h, w = int(Hh * vs), int(Ww * vs)
        if depth is None:
            planes, nf = depth_planes_first(batch['near_far'], D, h, w, rc.depth_inv[i])
        else:
            planes, nf = depth_planes_next(depth, std, nf, D, vs / rc.volume_scale[i - 1], rc.depth_inv[i - 1],
                                           rc.depth_inv[i])
        pm = proj_mats(exts, ixts, batch['tar_ext'], batch['tar_ixt'], rc.im_feat_scale[i], vs)
        vol = cost_volume_var(feats[f'level_{i}'], pm, planes)
        vol, logits = getattr(net, f'cost_reg_{i}')(vol)
        depth, std = depth_regression(logits, planes, rc.depth_inv[i])
        if not rc.render_if[i]:
            continue
        rays12 = build_rays(depth, std, nf, batch[f'rays_{i}'], rc.render_scale[i] / vs, rc.depth_inv[i])
        S = rc.num_samples[i]
        xyz, _, z = sample_along_depth(rays12, S, rc.depth_inv[i])
        H, W = int(Hh * rc.render_scale[i]), int(Ww * rc.render_scale[i])
        inv_scale = torch.tensor([W - 1, H - 1], dtype=torch.float32, device=xyz.device).unsqueeze(0).expand(B, -1)
        m = mask_viewport(xyz, exts, ixts, inv_scale).reshape(B, -1, S, 1) / S
        m = composite(m.repeat(1, 1, 1, 4), z, rc.white_bkgd)['rgb'].mean(-1)
        out[f'mask_level{i}'] = m.reshape(B, H, W)
    return out


def forward_view_selection(net, batch, rc):
    """reference lib/networks/boost_enerf/network.py:97-121 -> {"<scene>_<view>": [indices]}."""
    from oracle.mvsnerf_oracle import search_k_best_views
    N = batch['all_src_inps'].shape[1]
    per_triple = [calc_mask(net, t, batch, rc) for t in torch.combinations(torch.arange(N), 3)]
    result = {}
    for i in range(rc.num):
        if not rc.render_if[i]:
            continue
        picked = search_k_best_views([m[f'mask_level{i}'] for m in per_triple], rc.k_best)
        for j in range(len(batch['meta']['scene'])):
            result[f"{batch['meta']['scene'][j]}_{batch['meta']['tar_view'][j]}"] = picked
    return result
