"""Per-kernel timings + counters of the fusion path on one workload (run on a B200).

    python tests/tools/kernel_probe.py {tum|hires|c5|c5hires} [warm_frames] [steps] [--warm-l2]

Builds the map over `warm_frames`, then runs `steps` integrate + raycast frames with a 256 MiB L2 flush before each
and prints one JSON line: average per-kernel CUDA-event times (the library's own event pairs), per-frame counters
(visible / culled / updated / new / carved) and the integrate kernel's algorithmic GB/s.
RSF_LIBRARY selects a dev build of the library (A/B runs)."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
wl = sys.argv[1] if len(sys.argv) > 1 else "tum"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 40
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
flush_l2 = "--warm-l2" not in sys.argv
dev = torch.device("cuda", 0)
kw = {}
if wl == "tum":
    seq = synth.config2(n_frames=max(300, warm + steps))
elif wl == "hires":
    seq = synth.config3(n_frames=max(200, warm + steps), hires=True)
elif wl in ("c5", "c5hires"):
    seq = synth.config5(n_frames=warm + steps, **({"H": 968, "W": 1296} if wl == "c5hires" else {}))
    kw = dict(pool_blocks=1 << 21, table_slots_log2=23)
else:
    raise SystemExit("workload?")
H, W = seq.H, seq.W
grid = rsf.TSDFGrid(seq.voxel_size, seq.truncation, stream=torch.cuda.current_stream().cuda_stream, **kw)
cam = rsf.CameraParams(seq.K, H, W)
rmax = seq.max_depth * (1.25 if wl.startswith("c5") else 2.0)
out = {"rgba": torch.empty((H, W, 4), dtype=torch.uint8, device=dev),
       "normal": torch.empty((H, W, 4), dtype=torch.uint8, device=dev),
       "depth": torch.empty((H, W), dtype=torch.float32, device=dev)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if wl.startswith("c5"):
    frames = [synth.render_torch(seq.scene, seq.K, H, W, seq.poses[i], dev) for i in range(warm + steps)]
else:
    frames = [tuple(torch.from_numpy(a).to(dev) for a in seq.frame(i)) for i in range(warm + steps)]
for i in range(warm):
    grid.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    grid.RayCast(rmax, cam, seq.poses[i], out=out)
grid.synchronize()
grid.profile_enable(True)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
ctr = {k: 0 for k in ("visible_blocks", "culled_blocks", "updated_voxels", "new_blocks", "carved_blocks")}
for k in range(steps):
    i = warm + k
    if flush_l2:
        flush.zero_()
    ev[k][0].record()
    grid.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    grid.RayCast(rmax, cam, seq.poses[i], out=out)
    ev[k][1].record()
    st = grid.stats()          # blocking: keeps the counters per frame
    for n in ctr:
        ctr[n] += st[n]
torch.cuda.synchronize()
prof = grid.profile_read()
step_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
per = {k: v / steps for k, v in ctr.items()}
int_ms = prof["integrate"][0] / max(prof["integrate"][1], 1)
alg = 24.0 * per["updated_voxels"] + 12.0 * (per["visible_blocks"] - per["culled_blocks"]) + 15.0 * H * W
print(json.dumps({"workload": wl, "lib": os.environ.get("RSF_LIBRARY", "default"), "l2": "flushed" if flush_l2 else "warm",
                  "step_us": round(1e3 * step_ms, 1),
                  "kernel_us": {k: round(1e3 * v[0] / max(v[1], 1), 2) for k, v in prof.items()},
                  "per_frame": {k: round(v, 1) for k, v in per.items()},
                  "integrate_alg_MB": round(alg / 1e6, 2), "integrate_GBs": round(alg / (int_ms * 1e-3) / 1e9, 1),
                  "integrate_frac_of_6539": round(alg / (int_ms * 1e-3) / 1e9 / 6539.2, 3),
                  "active_blocks": st["active_blocks"], "table_keys": st["table_keys"], "checksum": float(out["depth"].sum())}))
