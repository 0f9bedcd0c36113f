"""Drop-in proof at the source level: the reference's OWN modules/tsdf_module.cc (and modules/tsdf_module.h) compile
UNMODIFIED against include/rsfusion/compat (an include overlay that answers  #include "utils/tsdf/voxel_tsdf.cuh"  with
rsfusion::TSDFGrid) and, on a GPU, produce the same render / dumps as the Python binding of the same library.

The build needs /root/reference (this container only; the binary travels to the GPU box with the snapshot); Eigen and
OpenCV are the stand-in headers of oracle/ref_shim, the GL image is a host-capturing stub (tests/cpp/gl_capture)."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "ref_module_main")
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources only exist in the development container")
def test_unmodified_tsdf_module_compiles_against_the_overlay():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "ra-slam_b200", "csrc")])
    subprocess.check_call(["make", "-B", "-C", os.path.dirname(EXE), "ref_module_main"])
    assert os.path.exists(EXE)
    # the reference's translation unit calls into the C ABI, nothing else
    und = subprocess.check_output(["nm", "-D", "--undefined-only", EXE], text=True)
    for sym in ("rsf_integrate", "rsf_raycast_internal_device", "rsf_gather_valid_semantic", "rsf_gather_valid_mesh", "rsf_gather_voxels"):
        assert sym in und, sym
    # ... and it was compiled from the reference tree, not from a copy
    assert not os.path.exists(os.path.join(ROOT, "tests", "cpp", "tsdf_module.cc"))


@pytest.mark.gpu
def test_reference_tsdf_system_runs_on_this_library(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("tests/cpp/ref_module_main was not built (needs /root/reference at build time)")
    rsf = importlib.import_module("ra-slam_b200")
    synth = importlib.import_module("ra-slam_b200.synth")
    seq = synth.config2(n_frames=300, H=120, W=160)
    n = 24
    g = rsf.TSDFGrid(seq.voxel_size, seq.truncation)
    fpath = tmp_path / "frames.bin"
    with open(fpath, "wb") as f:
        f.write(np.array([n, seq.H, seq.W], np.int32).tobytes())
        f.write(seq.K.as_array().tobytes())
        f.write(np.array([seq.voxel_size, seq.truncation, seq.max_depth], np.float32).tobytes())
        for i in range(n):
            fr = seq.frame(i)
            g.Integrate(*fr, seq.max_depth, seq.K, seq.poses[i])
            f.write(seq.poses[i].as_array().tobytes())
            for a in fr:
                f.write(np.ascontiguousarray(a).tobytes())
    out = subprocess.check_output([EXE, str(fpath), str(tmp_path / "out")], text=True, timeout=120)
    assert "loads 1 1" in out                      # TSDFSystem::Render handed both CUDA images to GLImage8UC4::LoadCuda
    ref = g.RayCast(seq.max_depth * 2, rsf.CameraParams(seq.K, seq.H, seq.W), seq.poses[n - 1])
    assert np.array_equal(np.fromfile(tmp_path / "out.rgba", np.uint8), ref["rgba"].ravel())
    assert np.array_equal(np.fromfile(tmp_path / "out.normal", np.uint8), ref["normal"].ravel())

    def rows_sorted(a):
        return a[np.lexsort(a.T[::-1])]
    rows = np.fromfile(tmp_path / "out.tsdf", np.float32).reshape(-1, 5)
    mine = g.GatherValidSemantic()
    assert rows.shape == mine.shape and np.array_equal(rows_sorted(rows), rows_sorted(mine))
    q = np.fromfile(tmp_path / "out.query", np.float32).reshape(-1, 4)
    qm = g.GatherVoxels(rsf.BoundingCube(0.0, 2.5, 0.0, 4.0, 0.0, 1.5))
    assert np.array_equal(rows_sorted(q), rows_sorted(qm))
    v, t, p = g.GatherValidMesh()
    assert np.fromfile(tmp_path / "out.vertices", np.float32).size == v.size
    assert np.fromfile(tmp_path / "out.indices", np.int32).size == t.size
    assert np.fromfile(tmp_path / "out.vprob", np.float32).size == p.size
