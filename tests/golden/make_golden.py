"""Regenerates tests/golden/oracle_small.npz from the CPU oracle (run from the repo root).

There is no runnable reference in this image, so this fixture pins the ORACLE against
silent change; it is not reference output.  hash_vectors.json holds the reference's own vectors.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rsfusion  # noqa: E402
from oracle_lib import OracleGrid  # noqa: E402

synth = importlib.import_module("ra-slam_b200.synth")
seq = synth.config2(n_frames=300, H=96, W=128)
g = OracleGrid(seq.voxel_size, seq.truncation, 2)
for i in range(12):
    g.Integrate(*seq.frame(i), seq.max_depth, seq.K, seq.poses[i])
c, t, w, p = g.ExportBlocks()
r = g.RayCast(seq.max_depth * 2, rsfusion.CameraParams(seq.K, seq.H, seq.W), seq.poses[11])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_small.npz"), coords=c,
                    rgbw_sum=w.astype(np.int64).sum(axis=(1, 2)), tsdf_sum=t.astype(np.float64).sum(axis=1),
                    prob_sum=p.astype(np.float64).sum(axis=1), raycast_hits=int((r["depth"] > 0).sum()),
                    raycast_depth=r["depth"])
print("blocks", len(c), "raycast hits", int((r["depth"] > 0).sum()))
