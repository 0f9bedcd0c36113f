"""Dev tool: characterise colour outliers between the product and the live reference at 1296x968 (run on a B200)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
import ref_lib  # noqa: E402
from oracle_lib import OracleGrid  # noqa: E402

rsf = importlib.import_module("ra-slam_b200")
synth = importlib.import_module("ra-slam_b200.synth")
seq = synth.config3(n_frames=60, hires=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
frames = [tuple(np.ascontiguousarray(t.cpu().numpy()) for t in
                synth.render_torch(seq.scene, seq.K, seq.H, seq.W, seq.poses[i], "cuda", depth_factor=seq.depth_factor)) for i in range(n)]
g = rsf.TSDFGrid(seq.voxel_size, seq.truncation)
r = ref_lib.RefGrid(seq.voxel_size, seq.truncation)
o = OracleGrid(seq.voxel_size, seq.truncation, 0)
for i in range(n):
    g.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    r.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
    o.Integrate(*frames[i], seq.max_depth, seq.K, seq.poses[i])
eg, er, eo = g.ExportBlocks(), r.ExportBlocks(), o.ExportBlocks()
kg, kr, ko = parity.keys_of(eg[0]), parity.keys_of(er[0]), parity.keys_of(eo[0])
print("blocks", len(kg), len(kr), len(ko), "product==oracle sets", np.array_equal(kg, ko))
m = parity.compare_blocks(eg, eo)
print("product vs oracle:", {k: m[k] for k in ("tsdf_bit_mismatch", "rgb_mismatch", "weight_mismatch", "prob_max_abs")})
both = np.isin(kg, kr)
eg = tuple(a[both] for a in eg)
kg = kg[both]
assert np.array_equal(kg, kr)
dc = np.abs(eg[2][..., :3].astype(int) - er[2][..., :3].astype(int)).max(axis=-1)
same = (np.abs(eg[1] - er[1]) <= 1e-4) & (eg[2][..., 3] == er[2][..., 3])
bad = (dc > 4) & same
print("colour outliers (>4) with equal tsdf/weight:", int(bad.sum()), "of", bad.size)
bi, vi = np.nonzero(bad)
print("weights of outliers (product):", np.bincount(eg[2][bi, vi, 3])[:42])
print("tsdf of outliers: min/max", eg[1][bi, vi].min(), eg[1][bi, vi].max())
for j in range(min(12, len(bi))):
    b, v = bi[j], vi[j]
    print("block", eg[0][b], "voxel", v, "product rgbw", eg[2][b, v], "reference rgbw", er[2][b, v], "tsdf", eg[1][b, v], er[1][b, v],
          "prob", eg[3][b, v], er[3][b, v])
# distribution over blocks
ub, cnt = np.unique(bi, return_counts=True)
print("outlier blocks:", len(ub), "max per block", cnt.max(), "voxel index pattern of the worst block:", np.sort(vi[bi == ub[np.argmax(cnt)]])[:40])
