"""Host-path timings on the TUM-shaped workload (run on a B200): frames/s of the blocking drop-in calls
(rsf_integrate + rsf_raycast, pinned host buffers) and of the pipelined calls (rsf_integrate_async + rsf_raycast_async +
rsf_wait, 3 frames in flight).  RSF_LIBRARY selects a dev build (A/B runs).  Dev tool; bench.py reports the same figures."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
W, K = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 100
seq = synth.config2(n_frames=max(300, W + K))
H, Wd = seq.H, seq.W
cam = rsf.CameraParams(seq.K, H, Wd)
pins = []


def pinned(a):
    pa = rsf.PinnedArray(a.shape, a.dtype)
    pa.array[...] = a
    pins.append(pa)
    return pa.array


frames = [tuple(pinned(a) for a in seq.frame(i)) for i in range(W + K)]
outs = [{k: pinned(np.zeros(s, d)) for k, s, d in (("rgba", (H, Wd, 4), np.uint8), ("normal", (H, Wd, 4), np.uint8), ("depth", (H, Wd), np.float32))}
        for _ in range(4)]
res = {"lib": os.environ.get("RSF_LIBRARY", "default").split("_")[-1].replace(".so", "")}
g = rsf.TSDFGrid(seq.voxel_size, seq.truncation)
for i in range(W):
    g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
g.synchronize()
t0 = time.perf_counter()
for k in range(K):
    i = W + k
    g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    g.RayCast(seq.max_depth * 2, cam, seq.poses[i], outputs=(), out=outs[0])
res["blocking_fps"] = K / (time.perf_counter() - t0)
res["checksum_blocking"] = float(outs[0]["depth"].sum())
g.close()
g = rsf.TSDFGrid(seq.voxel_size, seq.truncation)
for i in range(W):
    g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
g.synchronize()
t0 = time.perf_counter()
tickets = []
for k in range(K):
    i = W + k
    g.IntegrateAsync(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    tickets.append(g.RayCastAsync(seq.max_depth * 2, cam, seq.poses[i], outs[k % 4]))
    if len(tickets) > 3:
        g.wait(tickets.pop(0))
for t in tickets:
    g.wait(t)
g.synchronize()
res["pipelined_fps"] = K / (time.perf_counter() - t0)
res["checksum_pipelined"] = float(outs[(K - 1) % 4]["depth"].sum())
print(json.dumps(res))
