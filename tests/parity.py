"""Shared comparison helpers for the parity tests (SURVEY.md Appendix B tolerances)."""
import numpy as np


def keys_of(coords):
    c = np.asarray(coords).astype(np.int16).view(np.uint16).astype(np.uint64).reshape(-1, 3)
    return c[:, 0] | (c[:, 1] << np.uint64(16)) | (c[:, 2] << np.uint64(32))


def compare_blocks(a, b):
    """a, b = (coords, tsdf, rgbw, prob) sorted by key.  Returns a dict of parity metrics."""
    ka, kb = keys_of(a[0]), keys_of(b[0])
    sa, sb = set(ka.tolist()), set(kb.tolist())
    m = {"n_a": len(ka), "n_b": len(kb), "only_a": len(sa - sb), "only_b": len(sb - sa)}
    common = np.intersect1d(ka, kb)
    ia = np.searchsorted(ka, common)
    ib = np.searchsorted(kb, common)
    ta, tb = a[1][ia], b[1][ib]
    m["tsdf_max_abs"] = float(np.max(np.abs(ta - tb))) if len(common) else 0.0
    m["tsdf_bit_mismatch"] = int(np.sum(ta.view(np.uint32) != tb.view(np.uint32)))
    wa, wb = a[2][ia][..., 3].astype(np.int32), b[2][ib][..., 3].astype(np.int32)
    m["weight_mismatch"] = int(np.sum(wa != wb))
    ca, cb = a[2][ia][..., :3].astype(np.int32), b[2][ib][..., :3].astype(np.int32)
    m["rgb_max_abs"] = int(np.max(np.abs(ca - cb))) if len(common) else 0
    m["rgb_mismatch"] = int(np.sum(ca != cb))
    pa, pb = a[3][ia], b[3][ib]
    both_nan = np.isnan(pa) & np.isnan(pb)
    dp = np.where(both_nan, 0.0, np.abs(pa - pb))
    m["prob_max_abs"] = float(np.max(dp)) if len(common) else 0.0
    la, lb = pa > 0.5, pb > 0.5
    near_tie = np.abs(pa - 0.5) < 1e-6
    m["label_mismatch"] = int(np.sum((la != lb) & ~near_tie))
    m["label_mismatch_at_ties"] = int(np.sum((la != lb) & near_tie))
    m["n_common_voxels"] = int(len(common) * 512)
    return m


def assert_block_parity(m, *, prob_tol=1e-5, tsdf_tol=1e-4):
    assert m["only_a"] == 0 and m["only_b"] == 0, f"allocated block sets differ: {m}"
    assert m["tsdf_max_abs"] <= tsdf_tol, m
    assert m["weight_mismatch"] == 0, m
    assert m["rgb_max_abs"] <= 1, m
    assert m["prob_max_abs"] <= prob_tol, m
    assert m["label_mismatch"] == 0, m


def compare_raycast(a, b, voxel_size):
    """a, b = RayCast dicts.  Depth compared where both hit."""
    ha, hb = a["depth"] > 0, b["depth"] > 0
    both = ha & hb & np.isfinite(a["depth"]) & np.isfinite(b["depth"])
    m = {"hit_a": int(ha.sum()), "hit_b": int(hb.sum()), "hit_mask_mismatch": int((ha != hb).sum())}
    d = np.abs(a["depth"][both] - b["depth"][both]) / voxel_size
    m["depth_max_vox"] = float(d.max()) if d.size else 0.0
    m["depth_over_half_vox"] = int((d > 0.5).sum())
    m["prob_max_abs"] = float(np.abs(a["prob"][both] - b["prob"][both]).max()) if d.size else 0.0
    for k in ("rgba", "normal"):
        if k in a and k in b:
            dd = np.abs(a[k].astype(np.int32) - b[k].astype(np.int32))
            m[k + "_max_abs"] = int(dd.max())
            m[k + "_mismatch_px"] = int((dd.max(axis=-1) > 0).sum())
    return m
