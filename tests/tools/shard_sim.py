"""Per-rank kernel times of the spatially sharded scene, measured on ONE GPU: for each (world, rank) a grid with
that shard description integrates the same dense-walk building frames (configs[4], 1296x968 by default) and raycasts
them; broadcast / NVLink merge traffic are not part of this (they need the real ranks), the kernels are.

    python tests/tools/shard_sim.py "<world:rank,...>" [shift] [warm] [steps] [H W]

With a -DRSF_RAY_STATS build (RSF_LIBRARY=...) the raycast outputs are per-ray march counters and a summary is printed."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
pairs = [tuple(int(v) for v in p.split(":")) for p in (sys.argv[1] if len(sys.argv) > 1 else "1:0,2:0,8:0,8:3").split(",")]
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 4
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 12
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
H = int(sys.argv[5]) if len(sys.argv) > 5 else 968
W = int(sys.argv[6]) if len(sys.argv) > 6 else 1296
stats_build = "stats" in os.environ.get("RSF_LIBRARY", "")
dev = torch.device("cuda", 0)
seq = synth.config5(n_frames=600, H=H, W=W)
rmax = seq.max_depth * 1.25
cam = rsf.CameraParams(seq.K, H, W)
n = warm + steps
frames = [synth.render_torch(seq.scene, seq.K, H, W, seq.poses[i], dev) for i in range(n)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {"rgba": torch.empty((H, W, 4), dtype=torch.uint8, device=dev), "normal": torch.empty((H, W, 4), dtype=torch.uint8, device=dev),
       "depth": torch.empty((H, W), dtype=torch.float32, device=dev), "prob": torch.empty((H, W), dtype=torch.float32, device=dev)}
for world, rank in pairs:
    g = rsf.TSDFGrid(seq.voxel_size, seq.truncation, pool_blocks=1 << 21, table_slots_log2=23, shard_rank=rank, shard_world=world,
                     shard_shift=shift, stream=torch.cuda.current_stream().cuda_stream)
    for i in range(warm):
        g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
        g.RayCast(rmax, cam, seq.poses[i], out=out)
    g.synchronize()
    g.profile_enable(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        i = warm + k
        flush.zero_()
        ev[k][0].record()
        g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
        g.RayCast(rmax, cam, seq.poses[i], out=out)
        ev[k][1].record()
    torch.cuda.synchronize()
    prof = g.profile_read()
    st = g.stats()
    line = {"world": world, "rank": rank, "shift": shift, "step_us": round(1e3 * sum(a.elapsed_time(b) for a, b in ev) / steps, 1),
            "kernel_us": {k: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in prof.items()}, "active_blocks": st["active_blocks"],
            "visible": st["visible_blocks"], "culled": st["culled_blocks"]}
    if stats_build:
        real = out["depth"].cpu().numpy().astype(np.int64)
        fine = out["prob"].cpu().numpy().astype(np.int64)
        dbg = out["normal"].cpu().numpy().view(np.uint32)[..., 0]
        null, replay = (dbg & 0xffff).astype(np.int64), (dbg >> 16).astype(np.int64)
        found = out["rgba"].cpu().numpy()[..., 0] > 0
        line["rays"] = {"hit_frac": float(found.mean()), "real_mean": float(real.mean()), "real_hit_mean": float(real[found].mean()) if found.any() else 0,
                        "real_miss_mean": float(real[~found].mean()) if (~found).any() else 0, "fine_mean": float(fine.mean()),
                        "null_mean": float(null.mean()), "replayed_mean": float(replay.mean()), "real_p99": float(np.percentile(real, 99)),
                        "real_max": int(real.max())}
    print(json.dumps(line), flush=True)
    g.close()
