"""CUDA-graph replay of one boosted frame.

The reference's frame cannot be captured: it has ~100 implicit host syncs per frame.
Ours has none once the camera algebra is hoisted, so the whole frame — FPN, K cost-volume chains,
3-D CNNs, K3+K5, K4: ~300 launches — is captured once per (shape, selected triples) and replayed.
Per frame the host then only (1) copies the new batch into the static input buffers (for a host
batch this IS the H2D upload), (2) runs the ~1 KB camera algebra and uploads it, (3) replays.
Kernels read cameras from device memory (include/bmv.h conventions), which is what keeps the captured
graph valid across frames.
"""
from collections import OrderedDict

import torch

from .network import BoostEnerfNetwork, _combinations

_STATIC_KEYS = ("all_src_inps", "all_src_exts", "all_src_ixts", "tar_ext", "tar_ixt", "near_far")


# Network attributes that select kernels / precision: a captured graph bakes the routing in
_ROUTING_FLAGS = ("fold_bn", "channels_last", "fused_mlp", "half_feature_taps", "multi_chain_volume", "multi_chain_render", "mlp_engine",
                  "volume_range_scale", "host_camera_algebra", "overlap_fpn_topdown")


class FrameGraph:
    """max_entries bounds the number of captured graphs (each owns a full set of static buffers): least recently
    used entries are dropped, so a sequence whose selected triples change from view to view cannot grow without
    bound."""

    def __init__(self, net: BoostEnerfNetwork, max_entries=8, frame_fn=None):
        if not isinstance(net, BoostEnerfNetwork):
            raise TypeError("FrameGraph wraps a BoostEnerfNetwork")
        self.net = net
        self.max_entries = int(max_entries)
        # Network.forward leaves the LAST triple's views in the batch (batch['src_inps'] / ['src_exts'] / ['src_ixts'],
        # read by the reference's evaluator); a caller that does not read them can switch the three gathers off
        self.set_batch_views = True
        self._cache = OrderedDict()
        self._wtensors = None
        # frame_fn(static_inputs, camera, rays, triples, views_dev) -> output dict replaces the single-GPU frame body
        # (dist.make_sharded_graph: the multi-GPU frame with its NCCL collectives inside the captured graph)
        self.frame_fn = frame_fn

    def _key(self, batch, triples):
        """triples=None: the selection-agnostic key (graphs whose kernels read the view ids from device memory)."""
        """Everything a captured graph bakes in: shapes, the selected triples (kernel arguments), the weights (the graph
        holds pointers to PlanCache's folded copies and to the packed MLP / convolution weights, which are rebuilt when a
        parameter or buffer changes: same version-counter key as PlanCache) and the precision / kernel routing."""
        net = self.net
        # Walking the module tree costs ~0.3 ms per call; the tensor OBJECTS are collected once (again at every capture)
        # and only their (data_ptr, version) pairs are read per call.  load_state_dict, optimiser steps, .to() / .cuda()
        # all keep the Parameter objects; code that REPLACES a Parameter object must call invalidate().
        if self._wtensors is None:
            self._wtensors = list(net.parameters()) + list(net.buffers())
        return (tuple(batch["all_src_inps"].shape), None if triples is None else tuple(triples), bool(net.generate_rays),
                tuple(tuple(batch[f"rays_{i}"].shape) if f"rays_{i}" in batch else None for i in range(net.rc.num)),
                tuple((t.data_ptr(), t._version) for t in self._wtensors), id(net._plans),
                bool(torch.backends.cudnn.allow_tf32), bool(torch.backends.cuda.matmul.allow_tf32),
                tuple(getattr(net, f, None) for f in _ROUTING_FLAGS))

    def invalidate(self):
        """Drop every captured graph (after editing `param.data` directly, which bypasses the version counters)."""
        self._cache.clear()
        self._wtensors = None

    def close(self):
        """Release the captured graphs NOW.  Required before torch.distributed.destroy_process_group() when the graphs
        hold NCCL collectives (dist.make_sharded_graph): tearing the communicator down under a live graph hangs."""
        import gc
        self.invalidate()
        self.__dict__.pop("_rb", None)
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _triples(self, batch):
        net, rc = self.net, self.net.rc
        N = batch["all_src_inps"].shape[1]
        table = _combinations(N, rc.cost_volume_input_views)
        key = f"{batch['meta']['scene'][0]}_{batch['meta']['tar_view'][0]}"
        return [table[int(j)] for j in net.view_selection_outputs[key][:rc.k_best]]

    def _build(self, batch, triples):
        net, rc = self.net, self.net.rc
        dev = next(net.parameters()).device
        st = {k: torch.empty_like(batch[k], device=dev) for k in _STATIC_KEYS}
        gen_rays = net.generate_rays or any(f"rays_{i}" not in batch for i in range(rc.num))
        if not gen_rays:
            for i in range(rc.num):
                st[f"rays_{i}"] = torch.empty_like(batch[f"rays_{i}"], device=dev)
        N = st["all_src_inps"].shape[1]
        n_cam = rc.num * N * 12 + N * 3 + 3
        cam_dev = torch.zeros(n_cam, device=dev)
        cam_host = torch.zeros(n_cam).pin_memory()
        gen_dev = torch.zeros(rc.num * 12, device=dev, dtype=torch.float64)
        gen_host = torch.zeros(rc.num * 12, dtype=torch.float64).pin_memory()
        H, W = st["all_src_inps"].shape[-2:]
        K, I = len(triples), len(triples[0])
        views_dev = torch.zeros((K, I), device=dev, dtype=torch.int32)
        views_host = torch.zeros((K, I), dtype=torch.int32).pin_memory()
        entry = {"static": st, "views_dev": views_dev, "views_host": views_host, "triples": None, "cam_dev": cam_dev, "cam_host": cam_host, "gen_dev": gen_dev, "gen_host": gen_host,
                 "graph": None, "out": None}

        def frame():
            net._views_dev = views_dev                      # kernels read the view ids from here (replayable across selections)
            gens = net._raygen_views(gen_dev, (H, W)) if gen_rays else None
            camera = net._camera_views(cam_dev, st["all_src_exts"][0], st["all_src_ixts"][0]) + (gens,)
            rays = [None] * rc.num if gen_rays else [st[f"rays_{i}"][0] for i in range(rc.num)]
            if self.frame_fn is not None:
                try:
                    return self.frame_fn(st, camera, rays, triples, views_dev)
                finally:
                    net._views_dev = None
            lv = net._render_frame(st["all_src_inps"][0], st["all_src_exts"][0], st["all_src_ixts"][0], st["tar_ext"][0],
                                   st["tar_ixt"][0], st["near_far"][0], rays, triples, camera=camera)
            net._views_dev = None
            return net._assemble([lv])

        def body():
            out = frame()
            # batch['src_*'] of the LAST triple (see __call__), gathered INSIDE the graph: as three eager launches after
            # every replay they cost 36-43 us per frame (19 us of kernels + the gaps they open between consecutive graph
            # launches; tools/e2e_dissect.py)
            srcv = None
            if self.set_batch_views:
                last_idx = views_dev[-1].long()
                # (B = 1: advanced indexing of the squeezed tensor with the device-resident index — ATen's vectorised
                # gather, 7.6 us for the three images at C2, no host sync; torch.index_select takes its small-index
                # kernel, 21 us, along either dimension)
                srcv = {dst: st[src][0][last_idx].unsqueeze(0)
                        for src, dst in (("all_src_inps", "src_inps"), ("all_src_exts", "src_exts"), ("all_src_ixts", "src_ixts"))}
            return out, srcv

        self._load_views(entry, triples)
        self._load(entry, batch)
        entry["cam_loaded"] = False                     # the first real call always loads its own cameras
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                      # warm-up: plan caches, cuDNN handles, allocator pools
                body()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        net._baked_views = False
        try:
            with torch.cuda.graph(g), torch.no_grad():
                entry["out"], entry["src_views"] = body()
        finally:
            net._views_dev = None
        entry["graph"] = g
        # every launch read its view ids from views_dev: the graph serves ANY selection of K triples of this shape
        entry["agnostic"] = not net._baked_views
        return entry

    def _load_views(self, entry, triples):
        """Upload the frame's triples (K x 3 ints) when they changed; pinned staging buffer guarded by an event."""
        if entry["triples"] == triples:
            return
        if "views_ev" in entry:
            entry["views_ev"].synchronize()
        entry["views_host"].copy_(torch.tensor(triples, dtype=torch.int32))
        entry["views_dev"].copy_(entry["views_host"], non_blocking=True)
        entry.setdefault("views_ev", torch.cuda.Event()).record()
        entry["triples"] = list(triples)

    def prefetch(self, batch):
        """Start uploading the NEXT frame's host batch on a copy stream while the current frame renders.
        The following __call__ with the same batch object only does a device-to-device copy of the staged
        tensors before the replay.  No-op until the graph for this batch signature exists."""
        entry = self._cache.get(self._key(batch, None))
        if entry is None:
            entry = self._cache.get(self._key(batch, self._triples(batch)))
        if entry is None:
            return
        if "stage" not in entry:
            entry["stage"] = {k: torch.empty_like(t) for k, t in entry["static"].items()}
            entry["copy_stream"] = torch.cuda.Stream(device=entry["cam_dev"].device)
            entry["ev_staged"], entry["ev_consumed"] = torch.cuda.Event(), torch.cuda.Event()
            entry["ev_consumed"].record()
        cs = entry["copy_stream"]
        cs.wait_event(entry["ev_consumed"])                 # the previous staged frame has been copied out
        with torch.cuda.stream(cs):
            for k, t in entry["stage"].items():
                t.copy_(batch[k], non_blocking=True)
            entry["ev_staged"].record(cs)
        entry["staged_for"] = batch

    def read_back(self, out, host):
        """Pipelined read-back of a frame's results: `out` (the dict __call__ returned; its tensors are the graph's
        static outputs and are overwritten by the next replay) is copied device-to-device into a staging set on the
        current stream, and from there into the pinned `host` tensors (same keys) on a read-back stream — so the
        device-to-host transfer of frame i runs under the rendering of frame i+1, like prefetch() does for the upload.
        Call wait_read_back() (or synchronise) before touching `host`."""
        keys = [k for k in host if k in out]
        dev = out[keys[0]].device
        rb = self.__dict__.setdefault("_rb", {})
        if rb.get("sig") != tuple((k, tuple(out[k].shape)) for k in keys):
            rb.clear()
            rb["sig"] = tuple((k, tuple(out[k].shape)) for k in keys)
            rb["stage"] = {k: torch.empty_like(out[k]) for k in keys}
            rb["stream"] = torch.cuda.Stream(device=dev)
            rb["done"] = torch.cuda.Event()
            rb["done"].record()
        main = torch.cuda.current_stream()
        main.wait_event(rb["done"])                         # the previous frame's transfer has read the staging set
        for k in keys:
            rb["stage"][k].copy_(out[k], non_blocking=True)
        staged = main.record_event()
        rb["stream"].wait_event(staged)
        with torch.cuda.stream(rb["stream"]):
            for k in keys:
                host[k].copy_(rb["stage"][k], non_blocking=True)
            rb["done"].record(rb["stream"])

    def wait_read_back(self):
        """Make the current stream wait for the last read_back() transfer (then an event / synchronise on it covers
        the host copies too)."""
        rb = self.__dict__.get("_rb")
        if rb:
            torch.cuda.current_stream().wait_event(rb["done"])

    def _load(self, entry, batch, cameras_unchanged=False):
        net = self.net
        st = entry["static"]
        if entry.get("staged_for") is batch:                # uploaded by prefetch(): device-to-device only
            torch.cuda.current_stream().wait_event(entry["ev_staged"])
            for k, t in st.items():
                t.copy_(entry["stage"][k], non_blocking=True)
            entry["ev_consumed"].record()
            entry["staged_for"] = None
        else:
            for k, t in st.items():
                t.copy_(batch[k], non_blocking=True)
        N = st["all_src_inps"].shape[1]
        cams = [batch[k] for k in ("all_src_exts", "all_src_ixts", "tar_ext", "tar_ixt")]
        if all(c.device.type == "cpu" for c in cams):
            flat = torch.cat([c.reshape(-1) for c in cams])
        else:
            # Device-resident cameras are read back (one host sync per frame) unless the CALLER states that they are
            # the cameras of the previous call.  Tensor identity (data_ptr / _version) is not content identity: the
            # caching allocator hands the same block to the next frame's freshly uploaded cameras.
            if cameras_unchanged and entry.get("cam_loaded"):
                return
            flat = torch.cat([c.reshape(-1).to(st["near_far"].device) for c in cams]).cpu()
        # the pinned camera buffers are re-used every frame: wait until the previous frame's upload has executed
        # before overwriting them (the host may run a frame ahead of the GPU)
        if "cam_ev" in entry:
            entry["cam_ev"].synchronize()
        entry["cam_host"].copy_(net._camera_host(flat, N))
        entry["cam_dev"].copy_(entry["cam_host"], non_blocking=True)
        entry["gen_host"].copy_(net._raygen_host(flat, N))
        entry["gen_dev"].copy_(entry["gen_host"], non_blocking=True)
        entry.setdefault("cam_ev", torch.cuda.Event()).record()
        entry["cam_loaded"] = True

    def __call__(self, batch, cameras_unchanged=False):
        """batch: tensors on the GPU or in (pinned) host memory; B must be 1.  Returns the output dict;
        the tensors are the graph's static outputs and are overwritten by the next call.
        cameras_unchanged=True: the caller guarantees the four camera tensors hold the same values as in the previous
        call with this graph (skips the camera read-back of a device-resident batch)."""
        if self.net.training:
            raise RuntimeError("inference-only")
        if batch["all_src_inps"].shape[0] != 1:
            raise ValueError("FrameGraph renders one frame (B=1) per call")
        triples = self._triples(batch)
        key = self._key(batch, None)                         # selection-agnostic graph first
        entry = self._cache.get(key)
        if entry is None:
            key = self._key(batch, triples)
            entry = self._cache.get(key)
        if entry is None:
            while len(self._cache) >= max(1, self.max_entries):
                self._cache.popitem(last=False)              # least recently used graph + its static buffers
            self._wtensors = None                            # re-collect the weight tensors with every capture
            entry = self._build(batch, triples)
            key = self._key(batch, None if entry["agnostic"] else triples)
            self._cache[key] = entry
            cameras_unchanged = False
        else:
            self._cache.move_to_end(key)
        self._load_views(entry, triples)
        self._load(entry, batch, cameras_unchanged)
        entry["graph"].replay()
        # the reference leaves the LAST triple's views in the batch (evaluators read batch['src_inps'].shape;
        # reference lib/networks/boost_enerf/network.py:196-201): same contract as Network.forward
        # Gathered on the device from the uploaded copy (a host batch would pay an 18 MB CPU gather per frame) into
        # static tensors of the graph: like the returned outputs they are overwritten by the next call.
        # (index_select with a device-resident index: indexing with a Python list uploads the indices with a blocking
        # copy, i.e. one host sync per frame)
        if self.set_batch_views:
            if entry.get("src_views") is not None:           # gathered by the replay: static tensors, like `out`
                batch.update(entry["src_views"])
            else:                                            # graph captured with set_batch_views off
                last_idx = entry["views_dev"][-1].long()
                for src, dst in (("all_src_inps", "src_inps"), ("all_src_exts", "src_exts"), ("all_src_ixts", "src_ixts")):
                    batch[dst] = entry["static"][src].index_select(1, last_idx)
        return entry["out"]
