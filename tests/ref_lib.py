"""ctypes binding of oracle/_ref/libref_tsdf.so: RA-SLAM's OWN TSDFGrid (unmodified utils/tsdf/*.cu compiled
against oracle/ref_shim/, see oracle/Makefile target _ref and oracle/ref_harness.cu).  TEST INFRASTRUCTURE
ONLY: used by tests/ (parity against the real reference on the GPU) and bench.py --impl reference.  The
library is built in the development container (where /root/reference exists) and travels to the GPU box
as a prebuilt file; nothing here reads /root/reference at run time."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "libref_tsdf.so")
REFERENCE = "/root/reference"

_P = C.c_void_p
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def build(force: bool = False):
    """Compile the reference where it lies (only possible in the container that has /root/reference)."""
    if not os.path.isdir(REFERENCE):
        return LIB if os.path.exists(LIB) else None
    if force or not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "_ref"])
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} missing: run `make -C oracle _ref` where /root/reference exists")
    lib = C.CDLL(LIB)
    lib.ref_create.restype = _P
    lib.ref_create.argtypes = [C.c_float, C.c_float]
    lib.ref_destroy.argtypes = [_P]
    lib.ref_integrate.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P, _P]
    lib.ref_raycast.argtypes = [_P, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P]
    for n in ("ref_num_active_blocks", "ref_num_live_entries"):
        getattr(lib, n).restype = C.c_int64
        getattr(lib, n).argtypes = [_P]
    lib.ref_export_blocks.restype = C.c_int64
    lib.ref_export_blocks.argtypes = [_P, _P, _P, _P, _P, C.c_int64]
    for n in ("ref_gather_valid", "ref_gather_valid_semantic"):
        getattr(lib, n).restype = C.c_int64
        getattr(lib, n).argtypes = [_P, _P, C.c_int64]
    lib.ref_gather_voxels.restype = C.c_int64
    lib.ref_gather_voxels.argtypes = [_P, _P, _P, C.c_int64]
    lib.ref_gather_valid_mesh.argtypes = [_P, _P, C.c_int64, _P, _P, C.c_int64, _P, _P]
    lib.ref_last_cuda_error.restype = C.c_int
    _lib = lib
    return lib


def _sorted(coords, *arrs):
    c = coords.astype(np.int16).view(np.uint16).astype(np.uint64)
    order = np.argsort(c[:, 0] | (c[:, 1] << np.uint64(16)) | (c[:, 2] << np.uint64(32)), kind="stable")
    return (coords[order],) + tuple(a[order] for a in arrs)


class RefGrid:
    """Same method names as the product's TSDFGrid, backed by the reference's own CUDA TSDFGrid."""

    def __init__(self, voxel_size: float, truncation: float):
        self.lib = load()
        self.h = self.lib.ref_create(voxel_size, truncation)
        self.voxel_size, self.truncation = voxel_size, truncation

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_destroy(self.h)
            self.h = None

    def Integrate(self, rgb, depth, ht, lt, max_depth, intrinsics, cam_T_world):
        H, W = depth.shape
        rgb, depth, ht, lt = (np.ascontiguousarray(a) for a in (rgb, depth, ht, lt))
        assert rgb.dtype == np.uint8 and depth.dtype == ht.dtype == lt.dtype == np.float32
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        self.lib.ref_integrate(self.h, rgb.ctypes.data, depth.ctypes.data, ht.ctypes.data, lt.ctypes.data, H, W,
                               max_depth, K.ctypes.data, pose.ctypes.data)

    def integrate_raw(self, rgb_p, depth_p, ht_p, lt_p, H, W, max_depth, K, pose):
        """Pointer form for timing loops (no numpy work inside the timed region)."""
        self.lib.ref_integrate(self.h, rgb_p, depth_p, ht_p, lt_p, H, W, max_depth, K.ctypes.data, pose.ctypes.data)

    def raycast_raw(self, max_depth, H, W, K, pose, rgba_p=None, normal_p=None):
        self.lib.ref_raycast(self.h, max_depth, H, W, K.ctypes.data, pose.ctypes.data, rgba_p, normal_p)

    def RayCast(self, max_depth, virtual_cam, cam_T_world):
        H, W = virtual_cam.img_h, virtual_cam.img_w
        K, pose = virtual_cam.intrinsics.as_array(), cam_T_world.as_array()
        out = {"rgba": np.empty((H, W, 4), np.uint8), "normal": np.empty((H, W, 4), np.uint8)}
        self.lib.ref_raycast(self.h, max_depth, H, W, K.ctypes.data, pose.ctypes.data, out["rgba"].ctypes.data,
                             out["normal"].ctypes.data)
        return out

    def NumActiveBlock(self):
        return int(self.lib.ref_num_active_blocks(self.h))

    def NumLiveEntries(self):
        return int(self.lib.ref_num_live_entries(self.h))

    def ExportCoords(self):
        n = self.NumLiveEntries()
        coords = np.empty((max(n, 1), 3), np.int16)
        m = int(self.lib.ref_export_blocks(self.h, coords.ctypes.data, None, None, None, max(n, 1)))
        assert m == n
        return _sorted(coords[:n])[0]

    def ExportBlocks(self):
        n = self.NumLiveEntries()
        cap = max(n, 1)
        coords = np.empty((cap, 3), np.int16)
        tsdf = np.empty((cap, 512), np.float32)
        rgbw = np.empty((cap, 512, 4), np.uint8)
        prob = np.empty((cap, 512), np.float32)
        m = int(self.lib.ref_export_blocks(self.h, coords.ctypes.data, tsdf.ctypes.data, rgbw.ctypes.data,
                                           prob.ctypes.data, cap))
        assert m == n
        return _sorted(coords[:n], tsdf[:n], rgbw[:n], prob[:n])

    def _rows(self, fn, ncol, *pre):
        n = int(fn(self.h, *pre, None, 0))
        rows = np.empty((max(n, 1), ncol), np.float32)
        m = int(fn(self.h, *pre, rows.ctypes.data, n))
        assert m == n
        return rows[:n]

    def GatherValid(self):
        return self._rows(self.lib.ref_gather_valid, 4)

    def GatherValidSemantic(self):
        return self._rows(self.lib.ref_gather_valid_semantic, 5)

    def GatherVoxels(self, volumn):
        b = volumn.as_array()
        return self._rows(self.lib.ref_gather_voxels, 4, b.ctypes.data)

    def GatherValidMesh(self):
        nv, nt = C.c_int64(0), C.c_int64(0)
        self.lib.ref_gather_valid_mesh(self.h, None, 0, C.byref(nv), None, 0, C.byref(nt), None)
        v = np.empty((max(nv.value, 1), 3), np.float32)
        t = np.empty((max(nt.value, 1), 3), np.int32)
        p = np.empty(max(nv.value, 1), np.float32)
        nv2, nt2 = C.c_int64(0), C.c_int64(0)
        self.lib.ref_gather_valid_mesh(self.h, v.ctypes.data, nv.value, C.byref(nv2), t.ctypes.data, nt.value,
                                       C.byref(nt2), p.ctypes.data)
        n, m = min(nv.value, nv2.value), min(nt.value, nt2.value)
        return v[:n], t[:m], p[:n]
