"""The CPU oracle (and, on a GPU box, the product) against golden vectors produced by the REFERENCE ITSELF.

tests/golden/reference_small.npz was written by tests/golden/make_reference_golden.py on a B200 from RA-SLAM's
own TSDFGrid (unmodified utils/tsdf/*.cu, oracle/Makefile target _ref): 10 frames of the seeded config-2 sequence
at 96x72.  This is what pins the oracle: block sets, per-voxel weights and semantic labels must be identical,
TSDF / probability / colour within the north-star tolerance.

Known, bounded differences (see tests/test_gpu_vs_reference.py for the measurements on full-size inputs):
the reference build contracts FMAs (rounding-tie voxels may pick the neighbouring pixel), and its lossy
try-lock insert can allocate a block one frame late (28 of 2156 blocks in frame 0 of this fixture; they are one
observation short for ever and are excluded from the value comparisons, not from the set comparisons).
"""
import os

import numpy as np
import pytest

import parity
from oracle_lib import OracleGrid

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
TSDF_TOL = 1e-4


# reference_small: config 2 (TUM-shaped, 1 cm voxels, truncation 0.06); reference_config5: BASELINE.json configs[4]'s
# parameters (5 mm voxels, truncation 0.03, max_depth 8) on a window of the 1296x968 camera in the multi-room building
@pytest.fixture(scope="module", params=["small", "config5"])
def gold(request):
    path = os.path.join(GOLDEN_DIR, f"reference_{request.param}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} has not been generated (tests/golden/make_reference_golden.py {request.param} on a GPU box)")
    g = dict(np.load(path))
    g["which"] = request.param
    return g


def _sequence(synth, gold):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLDEN_DIR, "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    seq = mod.sequence(gold["which"])
    assert (seq.H, seq.W) == (int(gold["H"]), int(gold["W"]))
    return seq


def _check(grid, gold, seq, rsf, per_frame_sets=True):
    n = int(gold["n_frames"])
    counts = gold["counts"]
    frame_coords = np.split(gold["frame_coords"], np.cumsum(counts)[:-1])
    late = set()
    for i in range(n):
        grid.Integrate(*seq.frame(i), seq.max_depth, seq.K, seq.poses[i])
        if per_frame_sets:
            mine = set(parity.keys_of(grid.ExportBlocks()[0]).tolist())
            ref = set(parity.keys_of(frame_coords[i]).tolist())
            assert ref <= mine, f"frame {i}: the reference holds blocks this implementation lacks"
            late |= mine - ref
    assert len(late) <= 0.02 * counts[-1]       # blocks the reference's lossy insert delayed by a frame
    c, t, w, p = grid.ExportBlocks()
    assert np.array_equal(parity.keys_of(c), parity.keys_of(gold["coords"])), "final block-coordinate sets differ"
    clean = ~np.isin(parity.keys_of(c), np.array(sorted(late), dtype=np.uint64))
    # weights: identical apart from rounding-tie voxels
    wr = gold["weight"]
    assert (w[..., 3] != wr)[clean].mean() <= 2e-4, f"{(w[..., 3] != wr)[clean].sum()} weights differ"
    # semantic labels identical, ties excepted
    lab = np.unpackbits(gold["label"], axis=1).astype(bool)
    tie = np.unpackbits(gold["label_tie"], axis=1).astype(bool) | (np.abs(p - 0.5) <= 1e-6)
    assert (((p > 0.5) != lab) & ~tie)[clean].sum() == 0
    # full words on the stored subset
    pick = gold["pick"]
    pc = clean[pick][:, None]
    same_w = w[pick][..., 3] == wr[pick]
    dt = np.abs(t[pick] - gold["tsdf"])
    assert ((dt > TSDF_TOL) & pc).mean() <= 1e-3, f"{((dt > TSDF_TOL) & pc).sum()} of {dt.size} TSDF values outside 1e-4"
    ok = same_w & (dt <= TSDF_TOL) & pc
    assert np.nanmax(np.abs(p[pick] - gold["prob"])[ok]) <= 1e-4
    dc = np.abs(w[pick][..., :3].astype(int) - gold["rgb"].astype(int)).max(axis=-1)
    touched = (wr[pick] > 1) | (gold["tsdf"] != -1)
    # +-1 from u8 rounding; blocks on recycled pool slots start from stale colours in the reference
    assert (dc[ok & touched] > 1).mean() <= 0.02
    # raycast of the last frame: hit mask and colours
    ray = grid.RayCast(seq.max_depth * 2, rsf.CameraParams(seq.K, seq.H, seq.W), seq.poses[n - 1])
    ha, hb = ray["rgba"][..., 3] > 0, gold["ray_rgba"][..., 3] > 0
    assert hb.sum() > 1000
    assert (ha != hb).sum() <= 5e-3 * ha.size, f"hit masks differ in {(ha != hb).sum()} pixels"
    both = ha & hb
    d = np.abs(ray["rgba"].astype(int) - gold["ray_rgba"].astype(int))[both].max(axis=-1)
    dn = np.abs(ray["normal"].astype(int) - gold["ray_normal"].astype(int))[both].max(axis=-1)
    assert (d > 2).mean() <= 0.03 and (dn > 2).mean() <= 0.03, ((d > 2).mean(), (dn > 2).mean())
    assert len(grid.GatherValidSemantic()) == int(gold["gather_semantic_rows"])


def test_oracle_matches_reference_output(gold, synth, rsf):
    """CPU: the oracle restatement against the reference's own output."""
    seq = _sequence(synth, gold)
    _check(OracleGrid(seq.voxel_size, seq.truncation, 0), gold, seq, rsf)


@pytest.mark.gpu
def test_product_matches_reference_output(gold, synth, rsf):
    """GPU: the sm_100a path through the C ABI against the reference's own output (no reference library needed)."""
    seq = _sequence(synth, gold)
    _check(rsf.TSDFGrid(seq.voxel_size, seq.truncation), gold, seq, rsf)
