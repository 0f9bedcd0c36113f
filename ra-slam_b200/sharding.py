"""Spatial sharding of one large scene across the GPUs of a box (new, the reference
is single-GPU).

Data flow per frame, one process per GPU, ``torch.distributed`` (NCCL over NVLink; gloo in the CPU
tests):

1. ``broadcast_frame``: the ingest rank's RGB / depth / ht / lt images are packed into ONE byte
   buffer and broadcast with a single collective (4.6 MB at 640x480: latency-, not bandwidth-bound,
   so one call instead of four).
2. every rank integrates the frame into its own grid; the grid (``shard_rank`` / ``shard_world``)
   only allocates blocks it owns or that lie within the halo of an owned block, so integration
   work and voxel memory are partitioned.  Owner = mix(block_pos >> shift) % world with
   16^3-block super-bricks by default; the 1-block halo (integrated redundantly, bit-identical on
   every rank that holds it) keeps trilinear refinement, normals and the previous ray sample local.
3. every rank raycasts all pixels against its shard; a rank only reports a hit when it owns the
   block containing the zero crossing.  ``merge_raycast_min_depth`` then takes, per pixel, the hit
   with the smallest depth: one all-reduce(MIN) over int64 words ``depth_bits << 32 | payload``
   (positive IEEE floats order like unsigned integers).

Known, documented limit: the reference's ray march is stateful (a ray that grazed a surface keeps
the 0.3-voxel step for the rest of its iteration budget, voxel_tsdf.cu:361-367).  A rank that does
not hold the grazed surface keeps the coarse step and may report a hit behind it that the
single-GPU march never reaches; this affects the same ~0.2 % grazing pixels that the reference
renders black.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

_NO_HIT = 0x7F800000      # +inf bits: sorts after every finite positive depth


def frame_nbytes(H: int, W: int) -> int:
    return H * W * (3 + 4 + 4 + 4)


def pack_frame(rgb: torch.Tensor, depth: torch.Tensor, ht: torch.Tensor, lt: torch.Tensor, out: torch.Tensor | None = None):
    """[rgb u8 HxWx3 | depth f32 | ht f32 | lt f32] -> one contiguous uint8 buffer."""
    H, W = depth.shape
    if out is None:
        out = torch.empty(frame_nbytes(H, W), dtype=torch.uint8, device=depth.device)
    P = H * W
    out[:3 * P].copy_(rgb.reshape(-1))
    out[3 * P:7 * P].copy_(depth.reshape(-1).view(torch.uint8))
    out[7 * P:11 * P].copy_(ht.reshape(-1).view(torch.uint8))
    out[11 * P:15 * P].copy_(lt.reshape(-1).view(torch.uint8))
    return out


def unpack_frame(buf: torch.Tensor, H: int, W: int):
    P = H * W
    rgb = buf[:3 * P].view(H, W, 3)
    depth = buf[3 * P:7 * P].view(torch.float32).view(H, W)
    ht = buf[7 * P:11 * P].view(torch.float32).view(H, W)
    lt = buf[11 * P:15 * P].view(torch.float32).view(H, W)
    return rgb, depth, ht, lt


def broadcast_frame(frame, src: int = 0, group=None, buf: torch.Tensor | None = None):
    """Broadcast one frame in place with a single collective.

    `frame` = [rgb, depth, ht, lt] tensors (filled on `src`, preallocated elsewhere).
    Returns views into the packed buffer (valid until the next call that reuses `buf`).
    """
    rgb, depth, ht, lt = frame
    H, W = depth.shape
    if buf is None:
        buf = torch.empty(frame_nbytes(H, W), dtype=torch.uint8, device=depth.device)
    if dist.get_rank(group) == src:
        pack_frame(rgb, depth, ht, lt, out=buf)
    dist.broadcast(buf, src=src, group=group)
    out = unpack_frame(buf, H, W)
    for dst, s in zip(frame, out):
        if dst.data_ptr() != s.data_ptr():
            dst.copy_(s)
    return out


def _depth_key(depth: torch.Tensor) -> torch.Tensor:
    bits = depth.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    hit = depth > 0
    return torch.where(hit, bits, torch.full_like(bits, _NO_HIT))


def merge_raycast_min_depth(out: dict, group=None) -> dict:
    """Per-pixel min-depth merge of per-shard RayCast outputs (all ranks get the merged images).

    out: {'depth': HxW f32, and any of 'rgba' HxWx4 u8, 'normal' HxWx4 u8, 'prob' HxW f32}.
    One all-reduce(MIN) over an int64 tensor [n_payload, H, W] of depth_bits << 32 | payload.
    """
    depth = out["depth"]
    key = _depth_key(depth) << 32
    names, words = [], []
    for k in ("rgba", "normal", "prob"):
        if k in out:
            t = out[k].contiguous()
            if t.dtype == torch.uint8:
                payload = t.view(torch.int32).reshape(depth.shape).to(torch.int64) & 0xFFFFFFFF
            else:
                payload = t.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
            names.append(k)
            words.append(key | payload)
    if not words:
        names, words = [None], [key]
    packed = torch.stack(words)
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    kbits = (packed[0] >> 32) & 0xFFFFFFFF
    hit = kbits != _NO_HIT
    merged = {"depth": torch.where(hit, kbits.to(torch.int32).view(torch.float32), torch.zeros_like(depth))}
    for i, k in enumerate(names):
        if k is None:
            continue
        payload = torch.where(hit, packed[i] & 0xFFFFFFFF, torch.zeros_like(packed[i])).to(torch.int32)
        if out[k].dtype == torch.uint8:
            merged[k] = payload.view(torch.uint8).reshape(out[k].shape)
        else:
            merged[k] = payload.view(torch.float32).reshape(out[k].shape)
    return merged


class FusedMinDepthMerge:
    """Raycast + per-pixel min-depth merge as ONE kernel over NVLink peer memory (CUDA ranks of one node).

    The presenting rank (`root`) owns `n_planes * H * W` 64-bit words in torch symmetric memory; after the
    rendezvous every rank holds a device pointer to them (its own memory on the root, a peer mapping elsewhere).
    `raycast()` launches the grid's raycast kernel in merge mode: each hit a rank owns is a system-scope atomicMin of
    ``depth_bits << 32 | payload`` straight into the root's planes, so there is no intermediate image, no packing
    pass and no all-reduce -- only the pixels that hit travel, once.  A symmetric-memory barrier then orders all
    ranks' kernels before the root unpacks (and re-arms) the planes.  Replaces ``merge_raycast_min_depth`` (a
    7.4 MB NCCL all-reduce(MIN) at 640x480) on the data path; that function stays for gloo / CPU groups.
    """

    def __init__(self, H: int, W: int, device, n_planes: int = 3, root: int = 0, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.H, self.W, self.n_planes, self.root = H, W, n_planes, root
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.planes = symm_mem.empty(n_planes * H * W, dtype=torch.int64, device=device)
        self.planes.fill_(-1)                                   # all-ones = "no hit"
        self.hdl = symm_mem.rendezvous(self.planes, self.group)
        self.root_ptr = int(self.hdl.buffer_ptrs[root])
        torch.cuda.synchronize(device)
        self.hdl.barrier()

    def raycast(self, grid, max_depth, cam, pose, out: dict | None = None, rearm_barrier: bool = True):
        """All ranks call this.  Returns `out` filled on the root (None elsewhere).

        rearm_barrier=False skips the second barrier; only safe when something else orders the root's re-arming of
        the planes before any rank's next raycast -- e.g. the per-frame broadcast from the same root on the same
        stream, as in the integrate -> raycast loop of bench_sharded.py."""
        # the barrier below runs on torch's current stream: it only orders the raycast kernels if the grid runs there too
        gs = getattr(grid, "stream_handle", None)
        assert gs is not None and gs == torch.cuda.current_stream(self.planes.device).cuda_stream, \
            "FusedMinDepthMerge needs a grid created with stream=torch.cuda.current_stream().cuda_stream"
        grid.RayCastMerge(max_depth, cam, pose, self.root_ptr, self.n_planes)
        self.hdl.barrier()                                      # every rank's atomics have landed
        res = None
        if self.rank == self.root:
            res = grid.MergeFinalize(self.H, self.W, self.root_ptr, self.n_planes, out)
        if rearm_barrier:
            self.hdl.barrier()                                  # planes are re-armed before anyone writes again
        return res


class ShardedScene:
    """One large scene spatially sharded over the GPUs of a box, driven with ONE C-ABI call per frame per rank.

    Per frame: the ingest rank's packed frame reaches every rank (``feed``), then every rank calls
    ``rsf_shard_frame_device`` (``step``, C++): allocate (ownership first) | select -> integrate -> tile ranges -> raycast
    with the min-depth merge fused in (system-scope atomicMin into the presenting rank's planes over NVLink peer memory)
    -> arrival flag (release, system scope) in the presenting rank's memory; the presenting rank's stream then waits for all
    flags on the device and unpacks the planes.  No host-side barrier, no Python between the kernels.

    Frame transport, ``transport=``:
      "p2p"  (default) a binary tree of copy-engine pushes rooted at the ingest rank: a rank's copy engines push the frame
             from its receive buffer into its (at most two) children's receive buffers (symmetric memory, NVLink peer
             writes, one stream per child) followed by a 4-byte sequence word; forwarding and the compute stream wait
             for that word with a stream memory operation (rsf_stream_wait_flag_*).  No SM is involved, so the transfer
             of frame k + 1 really overlaps the kernels of frame k -- an NCCL kernel cannot: the fusion kernels are
             persistent grids that fill every SM, and the collective's CTAs only get in when a kernel ends (measured:
             90 us of every 0.3 ms step waiting for the broadcast at 8 ranks, 350 us without idle gaps between steps).
             A star (the ingest rank pushing to all seven peers, one after the other) is bound by its own egress: 7 x
             ~60 us per 18.8 MB frame made the pushes, not the kernels, the cycle time at 8 ranks.
      "nccl" one ncclBroadcast per frame on a side stream (the NCCL data plane; kept for comparison).
    Receive buffers and merge planes are double-buffered by frame parity: ``feed`` of frame k + 2 waits (stream event) for
    ``step`` of frame k on the ingest / presenting rank, which includes the device-side wait for every rank's arrival flag
    of frame k -- so by then every rank has consumed frame k's buffer and nobody can write a plane set before it has been
    re-armed (rsfusion.h, rsf_shard_link).  The ingest rank must be the presenting rank (`root`).
    """

    def __init__(self, voxel_size: float, truncation: float, H: int, W: int, device, *, root: int = 0, outputs=("rgba", "depth", "prob"),
                 group=None, pool_blocks: int = 0, table_slots_log2: int = 0, shard_shift: int = 4, shard_halo: int = 1,
                 transport: str = "p2p"):
        from . import TSDFGrid
        assert transport in ("p2p", "nccl")
        self.H, self.W, self.root, self.transport = H, W, root, transport
        n_planes = 3
        self.dist_on = dist.is_available() and dist.is_initialized()
        self.group = group if group is not None else (dist.group.WORLD if self.dist_on else None)
        self.rank = dist.get_rank(self.group) if self.dist_on else 0
        self.world = dist.get_world_size(self.group) if self.dist_on else 1
        self.device = torch.device(device)
        self.stream = torch.cuda.current_stream(self.device)
        self.grid = TSDFGrid(voxel_size, truncation, device=self.device.index, pool_blocks=pool_blocks,
                             table_slots_log2=table_slots_log2, shard_rank=self.rank, shard_world=self.world,
                             shard_shift=shard_shift, shard_halo=shard_halo, stream=self.stream.cuda_stream)
        self.nbytes = frame_nbytes(H, W)
        self.fed = [None, None]
        self.side = torch.cuda.Stream(self.device)
        self.ev_feed = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.n_fed = self.n_stepped = 0
        self.stall_events = None
        self.sym = None
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            P = H * W
            # symmetric allocation, in 64-bit words: [2 plane sets | arrival flags (64) | frame flags (64) | rx0 | rx1]
            w_planes = 2 * n_planes * P
            w_rx = (self.nbytes + 7) // 8
            self.sym = symm_mem.empty(w_planes + 128 + 2 * w_rx, dtype=torch.int64, device=self.device)
            self.sym[:w_planes].fill_(-1)
            self.sym[w_planes:].zero_()
            self.hdl = symm_mem.rendezvous(self.sym, self.group)
            base = int(self.hdl.buffer_ptrs[root])
            skip = (0 if "normal" in outputs else 2) | (0 if "prob" in outputs else 4)
            off_fflag = 8 * (w_planes + 64)
            off_rx = [8 * (w_planes + 128), 8 * (w_planes + 128 + w_rx)]
            mine = self.sym.view(torch.uint8)
            self.rx = [mine[off_rx[j]:off_rx[j] + self.nbytes] for j in range(2)]
            self.fflag_ptr = [int(self.hdl.buffer_ptrs[self.rank]) + off_fflag + 4 * j for j in range(2)]
            tr = None
            if transport == "p2p":
                me = (self.rank - root) % self.world         # position in the tree: children of node i are 2i + 1, 2i + 2
                kids = [(root + c) % self.world for c in (2 * me + 1, 2 * me + 2) if c < self.world]
                ptr = lambda r: int(self.hdl.buffer_ptrs[r])
                tr = {"frame_H": H, "frame_W": W, "rx": [ptr(self.rank) + off_rx[0], ptr(self.rank) + off_rx[1]],
                      "frame_flag": ptr(self.rank) + off_fflag,
                      "kids": [{"rx": [ptr(r) + off_rx[0], ptr(r) + off_rx[1]], "flag": ptr(r) + off_fflag} for r in kids]}
            self.grid.ShardSetLink(base, base + 8 * n_planes * P, base + 8 * w_planes, n_planes, root, H, W, skip, transport=tr)
            torch.cuda.synchronize(self.device)
            self.hdl.barrier()                               # everyone's planes are armed before the first frame
        else:
            self.rx = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(2)]

    def feed(self, packed_frame=None):
        """All ranks.  Start moving the next frame: `packed_frame` is the uint8 tensor in `pack_frame` layout on the ingest
        rank (None elsewhere).  Returns at once; the transfer runs on copy engines / a side stream."""
        j = self.n_fed & 1
        if self.rank == self.root:
            assert packed_frame is not None and packed_frame.numel() == self.nbytes
        if self.world > 1 and self.transport == "p2p":       # the library's copy-engine tree (rsf_shard_feed_device)
            self.fed[j] = packed_frame                       # keep the ingest rank's frame alive until it is consumed
            self.grid.ShardFeed(packed_frame.data_ptr() if self.rank == self.root else None)
            self.n_fed += 1
            return
        with torch.cuda.stream(self.side):
            if self.n_fed >= 2:
                self.side.wait_event(self.ev_done[j])        # frame n_fed - 2 is through on every rank: buffers j are free
            else:
                self.side.wait_stream(self.stream)
            if self.world > 1 and self.transport == "nccl":
                src = packed_frame if self.rank == self.root else self.rx[j]
                dist.broadcast(src, src=self.root, group=self.group, async_op=True).wait()
                self.fed[j] = src
            else:
                self.fed[j] = packed_frame if self.rank == self.root else self.rx[j]
            self.ev_feed[j].record(self.side)
        self.n_fed += 1

    def step(self, max_depth, intrinsics, cam_T_world, raycast_max_depth, virtual_cam, rc_cam_T_world, out: dict | None = None):
        """All ranks.  Fuse the oldest fed frame and render the virtual view; `out` (torch CUDA tensors) is filled on the
        presenting rank, asynchronously on the stream the scene was created on.  Returns `out` there, None elsewhere."""
        assert self.n_stepped < self.n_fed, "step() without a fed frame"
        assert torch.cuda.current_stream(self.device).cuda_stream == self.stream.cuda_stream, \
            "ShardedScene must be stepped on the stream it was created on"
        j = self.n_stepped & 1
        if self.world > 1 and self.transport == "p2p":       # rsf_shard_step_device: device-side wait for the frame + the fused frame
            o = out if (out is not None and self.rank == self.root) else {}
            ptrs = tuple(o[k].data_ptr() if k in o else None for k in ("rgba", "normal", "depth", "prob"))
            self.grid.ShardStep(max_depth, intrinsics.as_array(), cam_T_world.as_array(), raycast_max_depth,
                                virtual_cam.intrinsics.as_array(), rc_cam_T_world.as_array(), ptrs)
            self.n_stepped += 1
            return out if self.rank == self.root else None
        self.stream.wait_event(self.ev_feed[j])
        if self.stall_events is not None:                # measurement hook: when did the frame's data become available?
            e = torch.cuda.Event(enable_timing=True)
            e.record(self.stream)
            self.stall_events.append(e)
        fr = unpack_frame(self.fed[j], self.H, self.W)
        self.grid.ShardFrame(*fr, max_depth, intrinsics, cam_T_world, raycast_max_depth, virtual_cam, rc_cam_T_world,
                             out if self.rank == self.root else None)
        self.ev_done[j].record(self.stream)
        self.n_stepped += 1
        return out if self.rank == self.root else None

    def close(self):
        self.grid.close()


def gather_rows(rows: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank export rows (ownership is a partition, halo copies already dropped)."""
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, rows, group=group)
    return np.concatenate(parts, axis=0)
