#!/usr/bin/env python
"""bench_sharded.py -- one large scene spatially sharded over the GPUs of a box (BASELINE.json configs[4]: ~200 m^2
multi-room building, 5 mm voxels, truncation 0.03 m, max_depth 8 m).  `bench.py --gpus N` calls `run()` on every rank and
puts the result into its JSON line as "sharded_config5" (so the driver's 1 -> 8 scaling run carries the curve); this file
is also a stand-alone CLI:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_sharded.py --steps K --warmup W [--height 968 --width 1296] [--no-verify]

Per step (one frame), strong scaling (the same scene and frames for every N): the ingest rank (0) holds the packed frame in
device memory -> ONE NCCL broadcast -> every rank makes ONE C-ABI call (rsf_shard_frame_device: allocate | select ->
integrate -> tile ranges -> raycast with the per-pixel min-depth merge fused in over NVLink peer memory -> arrival flag ->
on rank 0: device-side wait for all flags + unpack).  N = 1 runs the same call on an unsharded grid (integrate + raycast).
Timing: CUDA events around each step on every rank, L2 flushed (256 MiB memset) between steps outside the events, max over
ranks.  Verification (default on, after the timed region): rank 0 replays the same frames into an unsharded grid and
compares the union of the ranks' owned block sets and the merged raycast of the last frame against it."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "configs[4] multi-room building 16x12 m, voxel 0.005, trunc 0.03, max_depth 8, {W}x{H}, dense walk (0.1 m / frame)"


def run(rank, local, world, steps=24, warmup=12, H=968, W=1296, pool_log2=21, shift=4, verify=True, breakdown=True, flush_l2=True, transport="p2p"):
    """All ranks call this (torch.distributed initialised when world > 1).  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    rsf = importlib.import_module("ra-slam_b200")
    synth = importlib.import_module("ra-slam_b200.synth")
    shard = importlib.import_module("ra-slam_b200.sharding")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    K, Wm = steps, warmup
    n = K + Wm
    # dense walk: the whole lawnmower path would be ~600 frames at 0.1 m per frame; the benchmark uses its first n frames
    seq = synth.config5(n_frames=600, H=H, W=W)
    rmax = seq.max_depth * 1.25
    cam = rsf.CameraParams(seq.K, H, W)
    scene = shard.ShardedScene(seq.voxel_size, seq.truncation, H, W, dev, pool_blocks=1 << pool_log2,
                               table_slots_log2=pool_log2 + 2, shard_shift=shift, transport=transport)
    grid = scene.grid
    # the ingest rank holds the frames in the wire format (one packed buffer per frame)
    packed = None
    if rank == 0:
        packed = [shard.pack_frame(*synth.render_torch(seq.scene, seq.K, H, W, seq.poses[i], dev)) for i in range(n)]
    out = {"rgba": torch.empty((H, W, 4), dtype=torch.uint8, device=dev), "depth": torch.empty((H, W), dtype=torch.float32, device=dev),
           "prob": torch.empty((H, W), dtype=torch.float32, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def feed(i):
        if i < n:
            scene.feed(packed[i] if rank == 0 else None)

    def step(i):
        scene.step(seq.max_depth, seq.K, seq.poses[i], rmax, cam, seq.poses[i], out)
        feed(i + 1)          # the next frame's broadcast overlaps this frame's kernels (side stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    feed(0)
    for i in range(Wm):
        step(i)
    grid.synchronize()
    tot0 = grid.totals()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    scene.stall_events = []
    import time as _time
    t_host0 = _time.perf_counter()
    for k in range(K):
        if flush_l2:
            flush.zero_()
        ev[k][0].record()
        step(Wm + k)
        ev[k][1].record()
    host_us = 1e6 * (_time.perf_counter() - t_host0) / K       # host time to ENQUEUE a step (no synchronisation inside the loop)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    # of each step: how long the stream sat waiting for the frame's broadcast to land (0 when the overlap works)
    stall_ms = sum(ev[k][0].elapsed_time(scene.stall_events[k]) for k in range(K)) / K if len(scene.stall_events) == K else 0.0
    scene.stall_events = None
    tot1 = grid.totals()
    st = grid.stats()
    merged_depth = out["depth"].clone() if rank == 0 else None
    owned = grid.ExportBlockCoords(owned_only=True) if verify else None      # before the breakdown re-integrates frames
    # per-kernel breakdown: the last frames once more with the library's per-kernel events and a broadcast timer
    # (the map is already built, so these are steady-state frames; not part of `value`)
    prof, bc_ms, Kd = None, 0.0, 0
    if breakdown:
        Kd = min(K, 8)
        grid.profile_enable(True)
        bev = []
        scratch = torch.empty(scene.nbytes, dtype=torch.uint8, device=dev)
        for k in range(Kd):
            i = Wm + K - Kd + k
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            buf = packed[i] if rank == 0 else scratch
            if world > 1:
                a.record()
                dist.broadcast(buf, src=0)
                b.record()
                bev.append((a, b))
            grid.ShardFrame(*shard.unpack_frame(buf, H, W), seq.max_depth, seq.K, seq.poses[i], rmax, cam, seq.poses[i],
                            out if rank == 0 else None)
        barrier()
        prof = grid.profile_read()
        grid.profile_enable(False)
        bc_ms = sum(a.elapsed_time(b) for a, b in bev) / max(len(bev), 1)
    vals = torch.tensor([ms, float(tot1["updated_voxels"] - tot0["updated_voxels"]), float(st["active_blocks"]),
                         float(tot1["visible_blocks"] - tot0["visible_blocks"])], dtype=torch.float64, device=dev)
    mx, sm = vals.clone(), vals.clone()
    mine = torch.tensor([(prof[k][0] / prof[k][1] if prof and prof[k][1] else 0.0) for k in ("allocate", "select", "integrate", "raycast", "merge")]
                        + [bc_ms, stall_ms, ms / K, host_us * 1e-3], dtype=torch.float64, device=dev)
    per_rank = [mine]
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(x) * 1e3, 1) for x in t.tolist()] for t in per_rank]
    ver = None
    if verify:
        parts = [owned]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, owned)
        if rank == 0:
            ref = rsf.TSDFGrid(seq.voxel_size, seq.truncation, device=local, pool_blocks=1 << pool_log2,
                               table_slots_log2=pool_log2 + 2, stream=torch.cuda.current_stream().cuda_stream)
            for i in range(n):
                ref.Integrate(*shard.unpack_frame(packed[i], H, W), seq.max_depth, seq.K, seq.poses[i])
            ref.synchronize()
            keys = rsf.block_keys(np.concatenate(parts))
            rk = rsf.block_keys(ref.ExportBlockCoords())
            r = ref.RayCast(rmax, cam, seq.poses[n - 1], outputs=("rgba", "depth", "prob"))
            md = merged_depth.cpu().numpy()
            valid = r["depth"] > 0
            both = valid & (md > 0)
            diff = np.abs(md[both] - r["depth"][both]) / seq.voxel_size
            ver = {"frames": n, "blocks_single": int(len(rk)), "blocks_union": int(len(keys)),
                   "duplicates": int(len(keys) - len(np.unique(keys))),
                   "block_set_equal": bool(np.array_equal(np.sort(keys), np.sort(rk))),
                   "hit_pixels_single": int(valid.sum()), "hit_mask_mismatch": int(((md > 0) != valid).sum()),
                   "depth_identical_frac": float((diff == 0).mean()) if diff.size else None,
                   "depth_max_diff_vox": float(diff.max()) if diff.size else None,
                   "depth_over_half_vox": int((diff > 0.5).sum())}
            ver["ok"] = bool(ver["block_set_equal"] and ver["duplicates"] == 0 and ver["hit_mask_mismatch"] <= 0.002 * H * W
                             and ver["depth_over_half_vox"] <= 0.002 * H * W)
            ref.close()
    scene.close()
    if rank != 0:
        return None
    ms_max = float(mx[0])
    names = ("allocate_us", "select_us", "integrate_us", "raycast_us", "merge_wait_unpack_us", "broadcast_serial_us",
             "timed_step_stall_on_broadcast_us", "timed_step_us", "host_enqueue_us_per_step")
    return {"metric": "frames/s integrate+raycast, one scene spatially sharded over N GPUs (strong scaling)",
            "value": K / (ms_max * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / K,
            "scaling": "strong", "workload": WORKLOAD.format(W=W, H=H),
            "config": {"shard_shift": shift, "halo_blocks": 1, "pool_blocks": 1 << pool_log2, "raycast_max_depth": rmax,
                       "l2": "flushed: 256 MiB memset between steps, outside the timed events" if flush_l2 else "not flushed",
                       "frame_transport": transport if world > 1 else None,
                       "per_step": (("rank 0's copy engines push the packed frame (%d bytes) + a sequence word into every rank's receive buffer over NVLink "
                                     "(frame k+1 moves while frame k is fused; ranks wait with a stream memory op)" if transport == "p2p" else
                                     "1 ncclBroadcast of the packed frame (%d bytes) on a side stream") % shard.frame_nbytes(H, W)) +
                                    " + 1 rsf_shard_frame_device per rank; merge = system-scope atomicMin into rank 0's planes over NVLink peer "
                                    "memory, arrival flags instead of a barrier"
                       if world > 1 else "1 rsf_shard_frame_device (unsharded grid: integrate + raycast)"},
            "nccl": {"backend": "nccl", "world_size": world, "collectives_per_step": 1 if world > 1 else 0} if world > 1 else None,
            "voxel_updates_per_s_all_ranks": float(sm[1]) / (ms_max * 1e-3),
            "active_blocks_sum_over_ranks": float(sm[2]), "active_blocks_max_rank": float(mx[2]),
            "visible_blocks_per_step_max_rank": float(mx[3]) / K,
            "per_rank_us": {"columns": names, "rows": per_rank, "frames": Kd,
                            "note": "library CUDA-event pairs per kernel (allocate and select overlap on two streams) + the NCCL broadcast, "
                                    "steady-state frames after the timed region"},
            "verify": ver}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--height", type=int, default=968)
    ap.add_argument("--width", type=int, default=1296)
    ap.add_argument("--pool-log2", type=int, default=21)
    ap.add_argument("--shift", type=int, default=4)
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--transport", choices=["p2p", "nccl"], default="p2p")
    ap.add_argument("--gpus", type=int, default=0)
    args = ap.parse_args()
    if "NCCL_DEBUG" not in os.environ:
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(rank, local, world, args.steps, args.warmup, args.height, args.width, args.pool_log2, args.shift, not args.no_verify,
              flush_l2=not args.no_flush, transport=args.transport)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
