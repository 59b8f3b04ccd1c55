"""Multi-GPU rendering of ONE frame (north_star: "rays (image row-tiles) and the K cost volumes shard
across the 8xB200 box, and the final frame is assembled with one NCCL all-gather over NVLink").

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  The reference has no multi-GPU inference at all
(SURVEY.md §2a); the partitioning follows the independence structure of the path
(reference lib/networks/boost_enerf/network.py:189-235):

  phase A  FeatureNet of the N source views      : replicated (default) — or views sharded over the ranks and the maps
                                                   all-gathered in fp16 (`shard_features=True`, TF32-class only)
  phase B  K cost-volume chains (K1, 3-D CNN, K2) : chain block [k0, k1) on rank r (each chain needs its WHOLE volume: the
                                                   U-Net's receptive field spans it)
  exchange #1 (all-to-all)                        : every chain owner sends every rank only the ROW SLAB of the regularised
                                                   volume and of the depth / std / near-far maps that the rank's rays touch
                                                   (1/G of the volume + 2 halo rows; fp16 volume in the TF32-class mode)
  phase C  K3+K5 (one launch for all K chains) + K4 for a contiguous row tile
  exchange #2 (all-gather)                        : the [rgb, depth, weights] row tiles = the frame, on every rank.

The whole frame, collectives included, is ONE CUDA graph per rank (`ShardedFrameGraph`): at this frame time (1-2 ms)
eager enqueueing is host-bound, and the per-frame host work is the same as for the single-GPU FrameGraph (upload the
inputs, ~1 KB of camera algebra, replay).

Everything that is pure host logic (partitioning, slab ranges, packing, re-assembly) is device-agnostic and covered by
world_size-2/3 gloo tests on CPU (tests/test_dist_gloo.py).
"""
import math

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------ partitioning
def owned_round_robin(n_items, world, rank):
    """Items i with i mod world == rank (source views of the sharded feature pyramid)."""
    return list(range(rank, n_items, world))


def chain_block(n_chains, world, rank):
    """Contiguous block [k0, k1) of chains owned by `rank`: ceil(K/G) per rank, trailing ranks may own none.
    (Contiguous, so a rank's chains are a slice of the K-stacked device tensors.)"""
    per = -(-n_chains // world)
    k0 = min(rank * per, n_chains)
    return k0, min(k0 + per, n_chains)


def row_tile(n_rows, world, rank):
    """Contiguous row range [r0, r1) of rank `rank`; tiles differ by at most one row."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def slab_rows(r0, r1, n_rows, grid_rows):
    """Rows [y0, y1) of a (grid_rows)-row map that rays of image rows [r0, r1) of an n_rows image read: the
    align_corners upsample of the depth maps (src = dst * (grid-1)/(n-1), taps floor(src), +1) and the trilinear fetch
    (same mapping), with one row of margin on each side for the fp32 rounding of `src`."""
    if r1 <= r0:
        return 0, 1
    s = (grid_rows - 1) / max(n_rows - 1, 1)
    y0 = max(0, int(math.floor(r0 * s)) - 1)
    y1 = min(grid_rows - 1, int(math.floor((r1 - 1) * s)) + 2)
    return y0, y1 + 1


def all_gather_ragged(local, counts, group=None):
    """local: (counts[rank], *item) tensor.  Returns the list over ranks of (counts[r], *item)
    tensors, using ONE all_gather_into_tensor on buffers padded to max(counts)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(counts) == world and local.shape[0] == counts[rank], (counts, local.shape, rank)
    item = tuple(local.shape[1:])
    cap = max(counts)
    if cap == 0:
        return [local.new_empty((0,) + item) for _ in range(world)]
    send = local.new_zeros((cap,) + item)
    send[:counts[rank]] = local
    recv = local.new_empty((world, cap) + item)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=group)
    return [recv[r, :counts[r]] for r in range(world)]


def interleave_round_robin(parts, n_items):
    """Inverse of owned_round_robin: parts[r][j] is item r + j*world; returns them in item order."""
    world = len(parts)
    out = [None] * n_items
    for r in range(world):
        for j, i in enumerate(range(r, n_items, world)):
            out[i] = parts[r][j]
    return out


def exchange_slabs(send_parts, counts_of, item_numel, like, group=None):
    """Exchange #1.  send_parts[q]: this rank's payload for rank q, a flat tensor of counts_of[rank] * item_numel[q]
    elements (its own chains' slabs cut for q's rows; empty when this rank owns no chain).  Returns the flat tensor
    received from every rank r (counts_of[r] * item_numel[rank] elements each), in rank order: ONE all_to_all_single
    with uneven splits — the chain owners are the only senders, everybody receives just its own rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_own = counts_of[rank]
    in_splits = [n_own * item_numel[q] for q in range(world)]
    out_splits = [counts_of[r] * item_numel[rank] for r in range(world)]
    send = torch.cat([p.reshape(-1) for p in send_parts]) if n_own else like.new_empty(0)
    recv = like.new_empty(sum(out_splits))
    dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
    return recv


# ------------------------------------------------------------------------------------------ sharded frame
class ShardedFrameRenderer:
    """Renders one boosted frame across the ranks of `group`.  `net` is a BoostEnerfNetwork replica
    (identical weights on every rank).  forward(batch) returns the same dict as net.forward(batch),
    complete on every rank."""

    def __init__(self, net, group=None, shard_features=False):
        self.net, self.group = net, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shard_features = bool(shard_features)

    # -- compute hooks (overridden by the CPU/gloo tests with stand-ins) ---------------------
    def compute_features(self, inps, views):
        """FeatureNet on the listed views -> dict level -> (len(views), C, h, w)."""
        if not len(views):
            return None
        if len(views) == inps.shape[0]:
            return self.net.forward_feat(inps)
        # round-robin ownership = a strided slice (indexing with a Python list would upload an index tensor: a host
        # sync, and illegal inside a CUDA-graph capture)
        step = views[1] - views[0] if len(views) > 1 else inps.shape[0]
        return self.net.forward_feat(inps[views[0]::step][:len(views)])

    def feature_shapes(self, inps):
        H, W = inps.shape[-2:]
        return {'level_0': (32, H // 4, W // 4), 'level_1': (16, H // 2, W // 2), 'level_2': (8, H, W)}

    def compute_chains(self, feats, projs, near_far, triples, H, W, views_dev=None):
        if not triples:
            return {}
        net = self.net
        keep = net._views_dev
        net._views_dev = views_dev
        try:
            return net._chain_levels(feats, projs, near_far, triples, H, W)
        finally:
            net._views_dev = keep

    def exchange_dtype(self, device):
        """fp16 volume slabs when the frame runs in the TF32-class mode (the regularised volume only feeds the trilinear
        fetch: 5e-4 relative on an input the MLP is insensitive to), fp32 in the strict mode.  The same on every rank,
        whether it owns a chain or not (sender and receiver must agree on the byte counts)."""
        return torch.float16 if (torch.backends.cudnn.allow_tf32 and device.type == 'cuda') else torch.float32

    def render_tile(self, level, feats, inps, vols, maps, y0, rays, cams, triples, H, W, ray_begin, n_rays, views_dev=None):
        """K3+K5 of all K chains + K4 for rays [ray_begin, ray_begin + n_rays).  vols (K,8,D,rows,wv) channels-last and
        maps (K,4,rows,wv) [depth, std, near, far] hold volume rows [y0, y0 + rows)."""
        from . import ops
        net, rc = self.net, self.net.rc
        S = rc.num_samples[level]
        K = vols.shape[0]
        rgb4 = feats.get('rgb_nhwc4')
        if rgb4 is None:
            rgb4 = inps.new_zeros((inps.shape[0], inps.shape[2], inps.shape[3], 4))
            rgb4[..., :3] = inps.permute(0, 2, 3, 1)
        rgb = rgb4.permute(0, 3, 1, 2)[:, :3]
        im_feat = feats[f'level_{rc.render_im_feat_level[level]}']
        if im_feat.dtype == torch.float16:
            im_feat = im_feat.float()
        hv = int(H * rc.volume_scale[level])
        dev = vols.device
        out = {'raw': torch.empty((K, n_rays, S, 4), device=dev), 'z_vals': torch.empty((K, n_rays, S), device=dev),
               'vis_mask': torch.empty((K, n_rays, S), device=dev)}
        ops.render_rays_multi(maps[:, 0], maps[:, 1], maps[:, 2:4], rays, H, W, rc.depth_inv[level], S, vols, im_feat, rgb, cams,
                              triples, net._packed_mlp(level, 'umma' if net.mlp_engine == 'umma' else 'mma'), ray_begin=ray_begin, n_rays=n_rays, out=out,
                              views_dev=views_dev, grid_rows=hv, vol_row0=y0, map_row0=y0)
        rgbo, depth, weights = ops.composite_blend(list(out['raw'].unbind(0)), list(out['vis_mask'].unbind(0)),
                                                   list(out['z_vals'].unbind(0)))
        return torch.cat([rgbo, depth[:, None], weights], dim=1)           # (n, 4+S)

    # -- exchanges ---------------------------------------------------------------------------
    def gather_features(self, inps, local, views_of):
        """all-gather F (shard_features): every level's maps of the views each rank computed, exchanged in fp16 / fp32 ->
        full (N,C,h,w) fp32 per level (channels-last physical layout, like FeatureNet emits them)."""
        N = inps.shape[0]
        counts = [len(v) for v in views_of]
        out = {}
        for name, (C, h, w) in self.feature_shapes(inps).items():
            # levels 0 / 1 only feed the cost volumes (TF32-class anyway): fp16 on the wire; level 2 is fetched by the
            # per-sample MLP, which amplifies input errors (5e-3 on rgb with fp16 maps, measured): fp32
            wire = torch.float32 if name == 'level_2' else torch.float16
            if local is None:
                mine = inps.new_empty((0, h, w, C), dtype=wire)
            else:
                mine = local[name].permute(0, 2, 3, 1).to(wire).contiguous()                 # (n,h,w,C) physical NHWC
            parts = all_gather_ragged(mine, counts, self.group)
            full = torch.empty((N, h, w, C), device=mine.device, dtype=torch.float32)
            for r, part in enumerate(parts):
                if part.shape[0]:
                    full[r::self.world] = part                      # round-robin ownership: view r + j*G
            out[name] = full.permute(0, 3, 1, 2)                    # logical NCHW, channels-last strides
        return out

    def exchange_chain_slabs(self, state, counts_of, rows_of, shape, dtype):
        """Exchange #1 for one rendered level.  state: this rank's chains (or None); rows_of[q] = (y0, y1) slab of rank
        q; shape = (Cv, D, hv, wv).  Returns vols (K,Cv,D,rows,wv) channels-last-3d fp32 and maps (K,4,rows,wv) fp32 for
        this rank's slab, chains in global order."""
        Cv, D, hv, wv = shape
        G, r = self.world, self.rank
        n_own = counts_of[r]
        K = sum(counts_of)
        like = state['feat_vol'] if state is not None else None
        dev = like.device if like is not None else shapes_device(self)
        vol_parts, map_parts = [], []
        if n_own:
            fv = state['feat_vol']                                  # (n_own, Cv, D, hv, wv), channels-last-3d
            maps = torch.stack([state['depth_all'], state['std_all'],
                                state['nf_all'][:, 0] if state['nf_all'].dim() == 4 else state['nf_all'][0].expand(n_own, -1, -1),
                                state['nf_all'][:, 1] if state['nf_all'].dim() == 4 else state['nf_all'][1].expand(n_own, -1, -1)], dim=1)
            for q in range(G):
                y0, y1 = rows_of[q]
                vol_parts.append(fv[:, :, :, y0:y1].permute(0, 2, 3, 4, 1).to(dtype).contiguous())   # (n,D,rows,wv,Cv)
                map_parts.append(maps[:, :, y0:y1].contiguous())
        y0, y1 = rows_of[r]
        rows = y1 - y0
        vol_n = [D * (rows_of[q][1] - rows_of[q][0]) * wv * Cv for q in range(G)]
        map_n = [4 * (rows_of[q][1] - rows_of[q][0]) * wv for q in range(G)]
        vrecv = exchange_slabs(vol_parts, counts_of, vol_n, torch.empty(0, device=dev, dtype=dtype), self.group)
        mrecv = exchange_slabs(map_parts, counts_of, map_n, torch.empty(0, device=dev, dtype=torch.float32), self.group)
        vols = vrecv.view(K, D, rows, wv, Cv)
        if vols.dtype != torch.float32:
            vols = vols.float()
        return vols.permute(0, 4, 1, 2, 3), mrecv.view(K, 4, rows, wv)

    def gather_frame(self, tile, rows_of, W):
        """Exchange #2 (the one north_star names): row tiles of [rgb, depth, weights] -> full frame."""
        counts = [(r1 - r0) * W for r0, r1 in rows_of]
        if len(set(counts)) == 1:                                   # equal tiles: gather straight into the frame buffer
            frame = tile.new_empty((self.world * counts[0], tile.shape[1]))
            dist.all_gather_into_tensor(frame.view(-1), tile.contiguous().view(-1), group=self.group)
            return frame
        return torch.cat(all_gather_ragged(tile, counts, self.group), dim=0)

    # -- the frame ---------------------------------------------------------------------------
    def frame(self, inps, exts, ixts, tar_ext, tar_ixt, near_far, rays_by_level, triples, camera=None, views_dev=None):
        """One frame, sharded.  Same arguments as BoostEnerfNetwork._render_frame; `camera` precomputed by the caller makes
        this capturable (ShardedFrameGraph).  Returns the output dict (complete on every rank)."""
        net, rc, G, r = self.net, self.net.rc, self.world, self.rank
        K = len(triples)
        N = inps.shape[0]
        H, W = inps.shape[-2:]
        if self.shard_features:
            views_of = [owned_round_robin(N, G, q) for q in range(G)]
            with net._stage('feature_net'):
                local = self.compute_features(inps, views_of[r])
            with net._stage('gather_features'):
                feats = self.gather_features(inps, local, views_of)
            if local is not None and 'rgb_nhwc4' in local and len(views_of[r]) == N:
                feats['rgb_nhwc4'] = local['rgb_nhwc4']
        else:
            with net._stage('feature_net'):
                feats = self.compute_features(inps, list(range(N)))
        need_gen = net.generate_rays or any(x is None for x in rays_by_level)
        with net._stage('camera'):
            cams, projs, gens = camera if camera is not None else net._camera_stage(exts, ixts, tar_ext, tar_ixt,
                                                                                    image_hw=(H, W) if need_gen else None)
        if need_gen:
            rays_by_level = list(gens)
        blocks = [chain_block(K, G, q) for q in range(G)]
        counts_of = [k1 - k0 for k0, k1 in blocks]
        k0, k1 = blocks[r]
        states_local = self.compute_chains(feats, projs, near_far, list(triples[k0:k1]), H, W,
                                           None if views_dev is None else views_dev[k0:k1])
        ret = {}
        owner0 = 0                                                  # chain 0 lives on rank 0
        for i in range(rc.num):
            if not rc.render_if[i]:
                continue
            hv, wv, D = int(H * rc.volume_scale[i]), int(W * rc.volume_scale[i]), rc.volume_planes[i]
            rs = rc.render_scale[i]
            Hr, Wr = int(H * rs), int(W * rs)
            tiles = [row_tile(Hr, G, q) for q in range(G)]
            rows_of = [slab_rows(t0, t1, Hr, hv) for t0, t1 in tiles]
            st = states_local.get(i)
            with net._stage('exchange_chains'):
                vols, maps = self.exchange_chain_slabs(st, counts_of, rows_of, (8, D, hv, wv), self.exchange_dtype(inps.device))
            t0, t1 = tiles[r]
            with net._stage(f'render_fused_l{i}'):
                tile = self.render_tile(i, feats, inps, vols, maps, rows_of[r][0], rays_by_level[i], cams, triples, H, W,
                                        t0 * Wr, (t1 - t0) * Wr, views_dev)
            with net._stage('gather_frame'):
                frame = self.gather_frame(tile, tiles, Wr)
                d0s0 = torch.empty((2, hv, wv), device=frame.device)
                if r == owner0:
                    d0s0[0], d0s0[1] = st['depth_all'][0], st['std_all'][0]
                dist.broadcast(d0s0, src=dist.get_global_rank(self.group, owner0) if self.group is not None else owner0,
                               group=self.group)
            ret.update({f'rgb_level{i}': frame[None, :, :3], f'depth_level{i}': frame[None, :, 3],
                        f'weights_level{i}': frame[None, :, 4:],
                        f'depth_mvs_level{i}': (1. / d0s0[0] if rc.depth_inv[i] else d0s0[0])[None],
                        f'std_level{i}': d0s0[1][None]})
        return ret

    def forward(self, batch):
        from .network import _combinations
        net, rc = self.net, self.net.rc
        net._check_mode(batch)
        inps_all = batch['all_src_inps']
        B, N = inps_all.shape[:2]
        if B != 1:
            raise ValueError("ShardedFrameRenderer renders one frame (B=1) per call")
        K, I = rc.k_best, rc.cost_volume_input_views
        table = _combinations(N, I)
        key = f"{batch['meta']['scene'][0]}_{batch['meta']['tar_view'][0]}"
        triples = [table[int(j)] for j in net.view_selection_outputs[key][:K]]
        with torch.no_grad():
            ret = self.frame(inps_all[0], batch['all_src_exts'][0], batch['all_src_ixts'][0], batch['tar_ext'][0],
                             batch['tar_ixt'][0], batch['near_far'][0],
                             [batch[f'rays_{i}'][0] if f'rays_{i}' in batch else None for i in range(rc.num)], triples)
            last = torch.tensor([table[int(net.view_selection_outputs[key][K - 1])]], device=inps_all.device)
            batch['src_inps'] = inps_all[:, last[0]]
            batch['src_exts'] = batch['all_src_exts'][:, last[0]]
            batch['src_ixts'] = batch['all_src_ixts'][:, last[0]]
        return ret


def shapes_device(renderer):
    return next(renderer.net.parameters()).device


def make_sharded_graph(net, group=None, shard_features=False, max_entries=4):
    """A FrameGraph whose captured body is the SHARDED frame: kernels and NCCL collectives of one frame as one CUDA graph
    per rank.  Every rank must call it with the same sequence of batch shapes (captures are collective)."""
    from .graph import FrameGraph
    sr = ShardedFrameRenderer(net, group, shard_features)

    def frame_fn(st, camera, rays, triples, views_dev):
        return sr.frame(st["all_src_inps"][0], st["all_src_exts"][0], st["all_src_ixts"][0], st["tar_ext"][0], st["tar_ixt"][0],
                        st["near_far"][0], rays, triples, camera=camera, views_dev=views_dev)
    fg = FrameGraph(net, max_entries=max_entries, frame_fn=frame_fn)
    fg.sharded = sr
    return fg
