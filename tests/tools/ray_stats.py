#!/usr/bin/env python
"""Per-ray march statistics of the raycast kernel (debug build -DRSF_RAY_STATS, build/stats/).  Dev tool."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("RSF_LIBRARY", os.path.join(ROOT, "build", "stats", "librsfusion_stats.so"))
rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
seq = synth.config2(n_frames=300, H=480, W=640)
g = rsf.TSDFGrid(seq.voxel_size, seq.truncation)
cam = rsf.CameraParams(seq.K, seq.H, seq.W)
res = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    g.Integrate(*seq.frame(i), seq.max_depth, seq.K, seq.poses[i])
    if i % 40 == 39:
        r = g.RayCast(seq.max_depth * 2, cam, seq.poses[i])
        real = r["depth"].astype(np.int64)
        fine = r["prob"].astype(np.int64)
        dbg = r["normal"].view(np.uint32)[..., 0]
        null, replay = (dbg & 0xffff).astype(np.int64), (dbg >> 16).astype(np.int64)
        found = r["rgba"][..., 0] > 0
        loww = r["rgba"][..., 1].astype(np.int64) | (r["rgba"][..., 2].astype(np.int64) << 8)
        long = real > 64
        longstats = {"n_long": int(long.sum()), "long_real_sum_frac": float(real[long].sum() / real.sum()),
                     "long_null_frac": float(null[long].sum() / max(real[long].sum(), 1)),
                     "long_loww_frac": float(loww[long].sum() / max(real[long].sum(), 1)),
                     "long_fine_frac": float(fine[long].sum() / max(real[long].sum(), 1)),
                     "long_hit_frac": float(found[long].mean()) if long.any() else 0,
                     "all_loww_mean": float(loww.mean())}
        # warp = 8x4 patch
        H, W = real.shape
        pr = real.reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
        res.append({"frame": i, "hit_frac": float(found.mean()), "real_mean": float(real.mean()), "real_p50": float(np.percentile(real, 50)),
                    "real_p90": float(np.percentile(real, 90)), "real_p99": float(np.percentile(real, 99)), "real_max": int(real.max()),
                    "fine_mean": float(fine.mean()), "null_mean": float(null.mean()), "replayed_mean": float(replay.mean()),
                    "warp_max_mean": float(pr.max(axis=1).mean()), "warp_mean_mean": float(pr.mean(axis=1).mean()),
                    "warp_max_p90": float(np.percentile(pr.max(axis=1), 90)), "warp_max_p99": float(np.percentile(pr.max(axis=1), 99)),
                    "warp_max_hist": np.histogram(pr.max(axis=1), bins=[0, 8, 16, 24, 32, 48, 64, 96, 128, 192, 268])[0].tolist(),
                    "miss_real_mean": float(real[~found].mean()) if (~found).any() else 0, "hit_real_mean": float(real[found].mean()), **longstats})
print(json.dumps(res, indent=1))
