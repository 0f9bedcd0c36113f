"""Small driver for ncu captures: builds a local map (device-resident frames), then runs a few
integrate + raycast steps.  Usage: python profiles/prof_driver.py [warm_frames] [steps] [H] [W]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 40
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H = int(sys.argv[3]) if len(sys.argv) > 3 else 480
W = int(sys.argv[4]) if len(sys.argv) > 4 else 640
seq = synth.config2(n_frames=300, H=H, W=W)
grid = rsf.TSDFGrid(seq.voxel_size, seq.truncation, stream=torch.cuda.current_stream().cuda_stream)
cam = rsf.CameraParams(seq.K, H, W)
out = {"rgba": torch.empty((H, W, 4), dtype=torch.uint8, device="cuda"),
       "normal": torch.empty((H, W, 4), dtype=torch.uint8, device="cuda"),
       "depth": torch.empty((H, W), dtype=torch.float32, device="cuda")}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frames = [tuple(torch.from_numpy(a).cuda() for a in seq.frame(i)) for i in range(warm + steps)]
for i in range(warm + steps):
    if i >= warm:
        flush.zero_()
    grid.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    grid.RayCast(seq.max_depth * 2, cam, seq.poses[i], out=out)
grid.synchronize()
print("ok", grid.stats(), float(out["depth"].sum()))
