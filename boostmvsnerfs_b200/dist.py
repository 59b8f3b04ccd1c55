"""Multi-GPU rendering of ONE frame (north_star: "rays (image row-tiles) and the K
cost volumes shard across the 8xB200 box, and the final frame is assembled with one NCCL
all-gather over NVLink").

One process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch).  The reference has no
multi-GPU inference at all; the partitioning below follows the independence
structure of the path:

  phase A  FeatureNet of the N source views      : views round-robin over ranks   -> all-gather F
  phase B  K cost-volume chains (K1, 3-D CNN, K2) : chain k on rank k mod G         -> all-gather #1
           (each chain needs its WHOLE volume: the U-Net's receptive field spans it)
  phase C  K3+MLP+K4 for a contiguous row tile    : rows [r*H/G, (r+1)*H/G) per rank -> all-gather #2
           of the rgb/depth/weights tiles = the frame, on every rank.

Payloads are small (C2: features 150 MB, volumes 4x33 MB, frame 10 MB), i.e. latency- not
bandwidth-bound on NVLink 5, so each exchange is ONE all_gather_into_tensor on a packed buffer.
Frame-level replication (different frames on different ranks, no collective) is the other mode and
lives in bench.py; this module is the single-frame latency mode.

Everything that is pure host logic (partitioning, packing, ragged gathers, re-assembly) is
device-agnostic and covered by world_size-2 gloo tests on CPU (tests/test_dist_gloo.py).
"""
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------ partitioning
def owned_round_robin(n_items, world, rank):
    """Items i with i mod world == rank (chains: north_star 'k mod G'; views likewise)."""
    return list(range(rank, n_items, world))


def row_tile(n_rows, world, rank):
    """Contiguous row range [r0, r1) of rank `rank`; tiles differ by at most one row."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


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


# ------------------------------------------------------------------------------------------ sharded frame
class ShardedFrameRenderer:
    """Renders one boosted frame across the ranks of `group`.  `net` is a BoostEnerfNetwork replica
    (identical weights on every rank).  forward(batch) returns the same dict as net.forward(batch),
    complete on every rank."""

    def __init__(self, net, group=None):
        self.net, self.group = net, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    # -- compute hooks (overridden by the CPU/gloo tests with stand-ins) ---------------------
    def compute_features(self, inps, views):
        """FeatureNet on the listed views -> dict level -> (len(views), C, h, w)."""
        if not views:
            return None
        feats = self.net.forward_feat(inps[views])
        # the exchange buffers are fp32: widen maps the single-GPU plan keeps in fp16
        return {k: (v.float() if torch.is_tensor(v) and v.dtype == torch.float16 else v) for k, v in feats.items()}

    def feature_shapes(self, inps):
        H, W = inps.shape[-2:]
        return {'level_0': (32, H // 4, W // 4), 'level_1': (16, H // 2, W // 2), 'level_2': (8, H, W)}

    def compute_chains(self, feats, projs, near_far, triples, H, W):
        return self.net._chain_levels(feats, projs, near_far, triples, H, W) if triples else {}

    def render_tile(self, level, feats, inps, state, rays, cams, triples, H, W, ray_begin, n_rays):
        from . import ops
        lv = self.net._render_level(level, feats, inps, state, rays, cams, triples, H, W, ray_begin, n_rays)
        rgb, depth, weights = ops.composite_blend(lv['raws'], lv['masks'], lv['zs'])
        return torch.cat([rgb, depth[:, None], weights], dim=1)           # (n, 4+S)

    # -- exchanges ---------------------------------------------------------------------------
    def gather_features(self, inps, local, views_of):
        """all-gather F: every level's maps of the views each rank computed -> full (N,C,h,w) per level
        (channels-last physical layout, like FeatureNet emits them)."""
        N = inps.shape[0]
        counts = [len(v) for v in views_of]
        out = {}
        for name, (C, h, w) in self.feature_shapes(inps).items():
            if local is None:
                mine = inps.new_empty((0, h, w, C))
            else:
                mine = local[name].permute(0, 2, 3, 1).contiguous()       # (n,h,w,C) physical NHWC
            parts = all_gather_ragged(mine, counts, self.group)
            full = torch.stack(interleave_round_robin(parts, N))          # (N,h,w,C)
            out[name] = full.permute(0, 3, 1, 2)                          # logical NCHW, channels-last strides
        return out

    def gather_chain_states(self, states_local, chains_of, K, shapes):
        """all-gather #1: regularised volume + depth/std/near_far maps of every chain, per rendered
        level.  shapes[level] = (Cv, D, h, w)."""
        counts = [len(c) for c in chains_of]
        out = {}
        for lvl, (Cv, D, h, w) in shapes.items():
            n_vol, n_map = Cv * D * h * w, h * w
            width = n_vol + 4 * n_map
            st = states_local.get(lvl)
            n_local = counts[self.rank]
            ref = st['feat_vol'] if st is not None else None
            if n_local == 0:
                dev = shapes_device(self)
                packed = torch.empty((0, width), device=dev)
            else:
                packed = torch.empty((n_local, width), device=ref.device)
                for j in range(n_local):
                    packed[j, :n_vol] = st['feat_vol'][j].permute(1, 2, 3, 0).reshape(-1)   # (D,h,w,C) order
                    packed[j, n_vol:n_vol + n_map] = st['depth'][j].reshape(-1)
                    packed[j, n_vol + n_map:n_vol + 2 * n_map] = st['std'][j].reshape(-1)
                    packed[j, n_vol + 2 * n_map:] = st['nf'][j].reshape(-1)
            parts = all_gather_ragged(packed, counts, self.group)
            rows = interleave_round_robin(parts, K)
            vols = torch.stack([r[:n_vol].view(D, h, w, Cv) for r in rows]).permute(0, 4, 1, 2, 3)
            out[lvl] = {'feat_vol': vols,
                        'depth': [r[n_vol:n_vol + n_map].view(h, w) for r in rows],
                        'std': [r[n_vol + n_map:n_vol + 2 * n_map].view(h, w) for r in rows],
                        'nf': [r[n_vol + 2 * n_map:].view(2, h, w) for r in rows]}
        return out

    def gather_frame(self, tile, rows_of, W):
        """all-gather #2 (the one north_star names): row tiles of [rgb, depth, weights] -> full frame."""
        counts = [(r1 - r0) * W for r0, r1 in rows_of]
        return torch.cat(all_gather_ragged(tile, counts, self.group), dim=0)

    # -- the frame ---------------------------------------------------------------------------
    def forward(self, batch):
        from .network import _combinations
        net, rc, G, r = self.net, self.net.rc, self.world, self.rank
        net._check_mode(batch)
        inps_all = batch['all_src_inps']
        B, N = inps_all.shape[:2]
        if B != 1:
            raise ValueError("ShardedFrameRenderer renders one frame (B=1) per call")
        K, I = rc.k_best, rc.cost_volume_input_views
        table = _combinations(N, I)
        key = f"{batch['meta']['scene'][0]}_{batch['meta']['tar_view'][0]}"
        triples = [table[int(j)] for j in net.view_selection_outputs[key][:K]]
        inps = inps_all[0]
        H, W = inps.shape[-2:]
        with torch.no_grad():
            views_of = [owned_round_robin(N, G, q) for q in range(G)]
            chains_of = [owned_round_robin(K, G, q) for q in range(G)]
            with net._stage('feature_net'):
                local = self.compute_features(inps, views_of[r])
            with net._stage('gather_features'):
                feats = self.gather_features(inps, local, views_of)
            with net._stage('camera'):
                cams, projs, _ = net._camera_stage(batch['all_src_exts'][0], batch['all_src_ixts'][0],
                                                   batch['tar_ext'][0], batch['tar_ixt'][0])
            states_local = self.compute_chains(feats, projs, batch['near_far'][0],
                                               [triples[k] for k in chains_of[r]], H, W)
            shapes = {i: (8, rc.volume_planes[i], int(H * rc.volume_scale[i]), int(W * rc.volume_scale[i]))
                      for i in range(rc.num) if rc.render_if[i]}
            with net._stage('gather_chains'):
                states = self.gather_chain_states(states_local, chains_of, K, shapes)
            ret = {}
            for i in shapes:
                rs = rc.render_scale[i]
                Hr, Wr = int(H * rs), int(W * rs)
                rows_of = [row_tile(Hr, G, q) for q in range(G)]
                r0, r1 = rows_of[r]
                tile = self.render_tile(i, feats, inps, states[i], batch[f'rays_{i}'][0], cams, triples, H, W,
                                        r0 * Wr, (r1 - r0) * Wr)
                with net._stage('gather_frame'):
                    frame = self.gather_frame(tile, rows_of, Wr)
                d0 = states[i]['depth'][0]
                ret.update({f'rgb_level{i}': frame[None, :, :3], f'depth_level{i}': frame[None, :, 3],
                            f'weights_level{i}': frame[None, :, 4:],
                            f'depth_mvs_level{i}': (1. / d0 if rc.depth_inv[i] else d0)[None],
                            f'std_level{i}': states[i]['std'][0][None]})
            last = torch.tensor([table[int(net.view_selection_outputs[key][K - 1])]], device=inps.device)
            batch['src_inps'] = inps_all[:, last[0]]
            batch['src_exts'] = batch['all_src_exts'][:, last[0]]
            batch['src_ixts'] = batch['all_src_ixts'][:, last[0]]
        return ret


def shapes_device(renderer):
    return next(renderer.net.parameters()).device
