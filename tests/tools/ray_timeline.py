#!/usr/bin/env python
"""Per-patch timeline of the raycast kernel (debug build -DRSF_RAY_TIMELINE in build/stats/): when each 8x4 patch started,
finished marching and finished, on which warp.  Dev tool; prints a summary of where the kernel's time goes."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("RSF_LIBRARY", os.path.join(ROOT, "build", "stats", "librsfusion_tl.so"))
rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
seq = synth.config2(n_frames=300, H=480, W=640)
g = rsf.TSDFGrid(seq.voxel_size, seq.truncation)
cam = rsf.CameraParams(seq.K, seq.H, seq.W)
out = []
for i in range(124):
    g.Integrate(*seq.frame(i), seq.max_depth, seq.K, seq.poses[i])
    r = g.RayCast(seq.max_depth * 2, cam, seq.poses[i]) if i >= 120 else None
    if r is None:
        continue
    w = r["normal"].view(np.uint32)[..., 0]
    rows = w[0::4]                                   # first row of every patch
    t0, tm, te, wid = rows[:, 0::8], rows[:, 1::8], rows[:, 2::8], rows[:, 3::8]
    base = int(t0.min())
    s = ((t0.astype(np.int64) - base) & 0xffffffff).ravel()
    m = ((tm.astype(np.int64) - base) & 0xffffffff).ravel()
    e = ((te.astype(np.int64) - base) & 0xffffffff).ravel()
    dur = e - s
    second = (wid.ravel() >> 16) & 1
    last = np.argsort(e)[-50:]
    span = float(e.max())
    # how many warps are busy over time (20 bins)
    bins = np.linspace(0, span, 21)
    busy = [int(((s < bins[k + 1]) & (e > bins[k])).sum()) for k in range(20)]
    out.append({"frame": i, "patches": int(dur.size), "kernel_span_us": span / 1e3,
                "patch_us_p50": float(np.percentile(dur, 50)) / 1e3, "patch_us_p90": float(np.percentile(dur, 90)) / 1e3,
                "patch_us_p99": float(np.percentile(dur, 99)) / 1e3, "patch_us_max": float(dur.max()) / 1e3,
                "march_share": float((m - s).sum() / max(dur.sum(), 1)),
                "pass1_patches": int((second == 0).sum()), "last_start_us": float(s.max()) / 1e3,
                "start_us_p50": float(np.percentile(s, 50)) / 1e3, "start_us_p90": float(np.percentile(s, 90)) / 1e3,
                "busy_warp_us_sum": float(dur.sum()) / 1e3,
                "dur_of_last_50_enders_us_mean": float(dur[last].mean()) / 1e3,
                "start_of_last_50_enders_us_mean": float(s[last].mean()) / 1e3,
                "pass_of_last_50_enders_second_frac": float(second[last].mean()),
                "patches_in_flight_per_5pct_of_span": busy})
print(json.dumps(out[-2:], indent=1))
