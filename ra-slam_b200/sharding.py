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
        grid.RayCastMerge(max_depth, cam, pose, self.root_ptr, self.n_planes)
        self.hdl.barrier()                                      # every rank's atomics have landed
        res = None
        if self.rank == self.root:
            res = grid.MergeFinalize(self.H, self.W, self.root_ptr, self.n_planes, out)
        if rearm_barrier:
            self.hdl.barrier()                                  # planes are re-armed before anyone writes again
        return res


def gather_rows(rows: np.ndarray, group=None) -> np.ndarray:
    """Concatenate per-rank export rows (ownership is a partition, halo copies already dropped)."""
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, rows, group=group)
    return np.concatenate(parts, axis=0)
