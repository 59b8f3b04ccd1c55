"""Full-frame diagnosis at C2: our strict path vs the reference op sequence on the SAME GPU, chain by chain
(depth / std per level, raw, visibility, z) and after the blend; counts how many rays differ and where.
Usage: python tools/diag_frame_precision.py [H W] [smooth]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from boostmvsnerfs_b200 import network, ops  # noqa: E402
from boostmvsnerfs_b200.config import RenderConfig  # noqa: E402
from boostmvsnerfs_b200.synth import batch_to, make_scene  # noqa: E402
from oracle import enerf_oracle as O  # noqa: E402


def stat(name, a, b, res, thr=1e-4):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    d = (a - b).abs()
    i = int(d.argmax())
    rng = float(b.abs().max())
    n_bad = int((d > thr * rng).sum())
    res[name] = {"max_abs": float(d[i]), "range": rng, "rel": float(d[i]) / max(rng, 1e-30), "n_bad": n_bad, "n": d.numel(), "argmax": i}
    print(f"{name:36s} max {float(d[i]):.3e} rel {float(d[i]) / max(rng, 1e-30):.2e} mean {float(d.mean()):.2e} n>{thr:g}*range {n_bad}/{d.numel()} argmax {i}")
    return d


def oracle_internals(net, batch, rc, k_best):
    inps = batch['all_src_inps']
    B, N = inps.shape[:2]
    I, K = rc.cost_volume_input_views, rc.k_best
    triples = O.view_triples(N, I, inps.device)[k_best]
    x = inps.view(B * N, *inps.shape[2:])
    f2, f1, f0 = net.feature_net(x)
    Hh, Ww = inps.shape[-2:]
    feats = {'level_2': f0.reshape(B, N, f0.shape[1], Hh, Ww), 'level_1': f1.reshape(B, N, f1.shape[1], Hh // 2, Ww // 2),
             'level_0': f2.reshape(B, N, f2.shape[1], Hh // 4, Ww // 4)}
    depth, std, nf = [None] * K, [None] * K, [None] * K
    bidx = torch.arange(B, device=inps.device).unsqueeze(-1).expand(-1, I)
    info = {'feats': feats, 'depth': {}, 'std': {}, 'vol': {}, 'cost': {}}
    per_k = []
    for i in range(rc.num):
        for k in range(K):
            vidx = triples[:, k]
            se, si = batch['all_src_exts'][bidx, vidx], batch['all_src_ixts'][bidx, vidx]
            D, vs = rc.volume_planes[i], rc.volume_scale[i]
            h, w = int(Hh * vs), int(Ww * vs)
            if depth[k] is None:
                planes, nf[k] = O.depth_planes_first(batch['near_far'], D, h, w, rc.depth_inv[i])
            else:
                planes, nf[k] = O.depth_planes_next(depth[k], std[k], nf[k], D, vs / rc.volume_scale[i - 1], rc.depth_inv[i - 1], rc.depth_inv[i])
            pm = O.proj_mats(se, si, batch['tar_ext'], batch['tar_ixt'], rc.im_feat_scale[i], vs)
            cost = O.cost_volume_var(feats[f'level_{i}'][bidx, vidx], pm, planes)
            vol, logits = getattr(net, f'cost_reg_{i}')(cost)
            depth[k], std[k] = O.depth_regression(logits, planes, rc.depth_inv[i])
            info['depth'][(i, k)], info['std'][(i, k)], info['vol'][(i, k)] = depth[k], std[k], vol
            if i == 0 and k == 0:
                info['cost'][(i, k)] = cost
            if not rc.render_if[i]:
                continue
            rays12 = O.build_rays(depth[k], std[k], nf[k], batch[f'rays_{i}'], rc.render_scale[i] / vs, rc.depth_inv[i])
            per_k.append(O.render_chain(rays12, vol, feats['level_2'][bidx, vidx], inps[bidx, vidx], se, si, batch['tar_ext'],
                                        getattr(net, f'nerf_{i}'), i, rc))
    raws = torch.stack([o['net_output'] for o in per_k], dim=1)
    masks_raw = torch.stack([o['mask'] for o in per_k], dim=1)
    masks = O.merge_masks(masks_raw, K)
    zs = torch.stack([o['z_vals'] for o in per_k], dim=1)
    out = O.composite_blend(raws, masks, zs, rc.white_bkgd)
    info.update(raws=raws, masks=masks_raw, zs=zs, out=out)
    return info


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (544, 960)
    smooth = len(sys.argv) > 3 and sys.argv[3] == "smooth"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    K, kb = 4, [0, 7, 12, 19]
    rc = RenderConfig.enerf_eval(K)
    torch.manual_seed(0)
    net = network.BoostEnerfNetwork(preprocess=True, rc=rc).eval().cuda()
    net.view_selection_outputs = {"synth_0": kb}
    batch = batch_to(make_scene(H=H, W=W, n_views=6, seed=0, smooth=smooth), "cuda")
    table = network._combinations(6, 3)
    triples = [table[j] for j in kb]
    res = {}
    with torch.no_grad():
        ref = oracle_internals(net, dict(batch), rc, torch.tensor([kb], device="cuda"))
        inps = batch["all_src_inps"][0]
        feats = net.forward_feat(inps)
        for lv in (0, 1, 2):
            stat(f"FPN level_{lv}", feats[f"level_{lv}"], ref['feats'][f"level_{lv}"][0], res)
        cams, projs, _ = net._camera_stage(batch["all_src_exts"][0], batch["all_src_ixts"][0], batch["tar_ext"][0], batch["tar_ixt"][0])
        # chain states, level by level (re-implemented from _chain_levels to see level 0 too)
        rcn = rc.with_(render_if=(True, True))
        old = net.rc
        object.__setattr__(net, 'rc', rcn) if False else None
        states = net._chain_levels(feats, projs, batch["near_far"][0], triples, H, W)
        st = states[1]
        for k in range(K):
            stat(f"chain {k} depth L1", st['depth'][k], ref['depth'][(1, k)][0], res)
            stat(f"chain {k} std L1", st['std'][k], ref['std'][(1, k)][0], res)
            stat(f"chain {k} feat_vol L1", st['feat_vol'][k], ref['vol'][(1, k)][0], res)
        lv = net._render_level(1, feats, inps, st, batch["rays_1"][0], cams, triples, H, W)
        for k in range(K):
            stat(f"chain {k} z", lv['zs'][k], ref['zs'][0, k], res)
            dm = stat(f"chain {k} vis mask", lv['masks'][k], ref['masks'][0, k].reshape(-1, 2), res)
            stat(f"chain {k} raw", lv['raws'][k], ref['raws'][0, k], res)
        rgb, depth, weights = ops.composite_blend(lv['raws'], lv['masks'], lv['zs'])
        d = stat("frame rgb", rgb, ref['out']['rgb'][0], res)
        stat("frame depth", depth, ref['out']['depth'][0], res)
        stat("frame weights", weights, ref['out']['weights'][0], res)
        # blend kernel on the REFERENCE's own per-chain tensors
        r2, d2, w2 = ops.composite_blend([ref['raws'][0, k].contiguous() for k in range(K)],
                                         [ref['masks'][0, k].reshape(-1, 2).contiguous() for k in range(K)],
                                         [ref['zs'][0, k].contiguous() for k in range(K)])
        stat("K4 on reference inputs: rgb", r2, ref['out']['rgb'][0], res)
        stat("K4 on reference inputs: depth", d2, ref['out']['depth'][0], res)
        stat("K4 on reference inputs: weights", w2, ref['out']['weights'][0], res)
        # where are the bad rays?
        per_ray = d.reshape(-1, 3).max(-1).values
        bad = (per_ray > 1e-4 * float(ref['out']['rgb'].abs().max())).nonzero().reshape(-1)
        print("bad rays:", bad.numel(), "of", per_ray.numel())
        if bad.numel():
            ys, xs = bad // W, bad % W
            print("  x range", int(xs.min()), int(xs.max()), "y range", int(ys.min()), int(ys.max()))
            hist = torch.histc(per_ray[bad].log10(), bins=8, min=-5, max=-1)
            print("  log10(err) histogram [-5..-1]:", [int(v) for v in hist])
            # do the bad rays coincide with visibility-count differences?
            any_flip = torch.zeros_like(per_ray, dtype=torch.bool)
            for k in range(K):
                any_flip |= ((lv['masks'][k] - ref['masks'][0, k].reshape(-1, 2)).abs() > 0).any(-1)
            print("  rays with a flipped visibility count:", int(any_flip.sum()), " bad rays among them:", int(any_flip[bad].sum()))
            res["bad_rays"] = {"n": int(bad.numel()), "flipped": int(any_flip.sum()), "bad_and_flipped": int(any_flip[bad].sum())}
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"diag_frame_precision_{W}x{H}{'_smooth' if smooth else ''}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
