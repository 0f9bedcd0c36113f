"""ctypes binding of the CPU oracle (oracle/rsf_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_P = C.c_void_p
_lib = None


def build(force: bool = False):
    src = os.path.join(ORACLE_DIR, "rsf_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB)
    lib.ora_create.restype = _P
    lib.ora_create.argtypes = [C.c_float, C.c_float, C.c_int]
    lib.ora_destroy.argtypes = [_P]
    lib.ora_set_shard.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ora_hash.restype = C.c_uint32
    lib.ora_hash.argtypes = [C.c_int16, C.c_int16, C.c_int16, C.c_uint32]
    lib.ora_shard_owner.restype = C.c_uint32
    lib.ora_shard_owner.argtypes = [C.c_int16, C.c_int16, C.c_int16, C.c_int, C.c_int]
    lib.ora_integrate.argtypes = [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P, _P]
    lib.ora_raycast.argtypes = [_P, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]
    lib.ora_stats.argtypes = [_P, _P]
    lib.ora_num_blocks.restype = C.c_int64
    lib.ora_num_blocks.argtypes = [_P]
    lib.ora_allocate_block.argtypes = [_P, C.c_int16, C.c_int16, C.c_int16]
    lib.ora_delete_block.argtypes = [_P, C.c_int16, C.c_int16, C.c_int16]
    lib.ora_retrieve.restype = C.c_int
    lib.ora_retrieve.argtypes = [_P, C.c_int16, C.c_int16, C.c_int16, _P]
    lib.ora_write_voxel.restype = C.c_int
    lib.ora_write_voxel.argtypes = [_P, C.c_int16, C.c_int16, C.c_int16, C.c_float, C.c_float, _P]
    lib.ora_trilinear.restype = C.c_float
    lib.ora_trilinear.argtypes = [_P, C.c_float, C.c_float, C.c_float]
    lib.ora_export_blocks.restype = C.c_int64
    lib.ora_export_blocks.argtypes = [_P, C.c_int, _P, _P, _P, _P]
    for n in ("ora_gather_valid", "ora_gather_valid_semantic"):
        getattr(lib, n).restype = C.c_int64
        getattr(lib, n).argtypes = [_P, _P]
    lib.ora_integrate_classes.argtypes = [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]
    lib.ora_gather_valid_classes.restype = C.c_int64
    lib.ora_gather_valid_classes.argtypes = [_P, _P]
    lib.ora_gather_valid_mesh.restype = C.c_int64
    lib.ora_gather_valid_mesh.argtypes = [_P, _P, _P, _P, _P]
    lib.ora_gather_voxels.restype = C.c_int64
    lib.ora_gather_voxels.argtypes = [_P, _P, _P]
    _lib = lib
    return lib


class OracleGrid:
    """Same method names as the product's TSDFGrid, backed by the CPU restatement."""

    def __init__(self, voxel_size: float, truncation: float, threads: int = 0):
        self.lib = load()
        self.h = self.lib.ora_create(voxel_size, truncation, threads)
        self.voxel_size, self.truncation = voxel_size, truncation

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ora_destroy(self.h)
            self.h = None

    def set_shard(self, rank, world, shift=4, halo=1):
        self.lib.ora_set_shard(self.h, rank, world, shift, halo)

    def Integrate(self, rgb, depth, ht, lt, max_depth, intrinsics, cam_T_world):
        H, W = depth.shape
        rgb, depth, ht, lt = (np.ascontiguousarray(a) for a in (rgb, depth, ht, lt))
        assert rgb.dtype == np.uint8 and depth.dtype == ht.dtype == lt.dtype == np.float32
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        self.lib.ora_integrate(self.h, rgb.ctypes.data, depth.ctypes.data, ht.ctypes.data, lt.ctypes.data, H, W,
                               max_depth, K.ctypes.data, pose.ctypes.data)

    def RayCast(self, max_depth, virtual_cam, cam_T_world):
        H, W = virtual_cam.img_h, virtual_cam.img_w
        K, pose = virtual_cam.intrinsics.as_array(), cam_T_world.as_array()
        out = {"rgba": np.empty((H, W, 4), np.uint8), "normal": np.empty((H, W, 4), np.uint8),
               "depth": np.empty((H, W), np.float32), "prob": np.empty((H, W), np.float32)}
        self.lib.ora_raycast(self.h, max_depth, H, W, K.ctypes.data, pose.ctypes.data, out["rgba"].ctypes.data,
                             out["normal"].ctypes.data, out["depth"].ctypes.data, out["prob"].ctypes.data)
        return out

    def stats(self):
        s = np.zeros(4, np.int64)
        self.lib.ora_stats(self.h, s.ctypes.data)
        return {"new_blocks": int(s[0]), "visible_blocks": int(s[1]), "updated_voxels": int(s[2]),
                "carved_blocks": int(s[3]), "active_blocks": self.NumActiveBlock()}

    def NumActiveBlock(self):
        return int(self.lib.ora_num_blocks(self.h))

    def ExportBlocks(self, owned_only=False):
        n = int(self.lib.ora_export_blocks(self.h, int(owned_only), None, None, None, None))
        coords = np.empty((n, 3), np.int16)
        tsdf = np.empty((n, 512), np.float32)
        rgbw = np.empty((n, 512, 4), np.uint8)
        prob = np.empty((n, 512), np.float32)
        if n:
            self.lib.ora_export_blocks(self.h, int(owned_only), coords.ctypes.data, tsdf.ctypes.data,
                                       rgbw.ctypes.data, prob.ctypes.data)
        return coords, tsdf, rgbw, prob

    def _rows(self, fn, ncol, *pre):
        n = int(fn(self.h, *pre, None))
        rows = np.empty((n, ncol), np.float32)
        if n:
            fn(self.h, *pre, rows.ctypes.data)
        return rows

    def GatherValid(self):
        return self._rows(self.lib.ora_gather_valid, 4)

    def GatherValidSemantic(self):
        return self._rows(self.lib.ora_gather_valid_semantic, 5)

    def GatherVoxels(self, volumn):
        b = volumn.as_array()
        return self._rows(self.lib.ora_gather_voxels, 4, b.ctypes.data)

    def IntegrateClasses(self, rgb, depth, class_prob, max_depth, intrinsics, cam_T_world):
        H, W = depth.shape
        rgb, depth, class_prob = (np.ascontiguousarray(a) for a in (rgb, depth, class_prob))
        assert class_prob.dtype == np.float32 and class_prob.shape[1:] == (H, W)
        self.num_classes = class_prob.shape[0]
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        self.lib.ora_integrate_classes(self.h, rgb.ctypes.data, depth.ctypes.data, class_prob.ctypes.data, self.num_classes, H, W,
                                       max_depth, K.ctypes.data, pose.ctypes.data)

    def GatherValidClasses(self):
        return self._rows(self.lib.ora_gather_valid_classes, 4 + self.num_classes)

    def GatherValidMesh(self):
        nt = C.c_int64(0)
        nv = int(self.lib.ora_gather_valid_mesh(self.h, None, None, None, C.byref(nt)))
        v = np.empty((nv, 3), np.float32)
        p = np.empty(nv, np.float32)
        t = np.empty((nt.value, 3), np.int32)
        if nv or nt.value:
            self.lib.ora_gather_valid_mesh(self.h, v.ctypes.data, p.ctypes.data, t.ctypes.data, C.byref(nt))
        return v, t, p

    # table-level
    def AllocateBlocks(self, coords):
        for x, y, z in np.asarray(coords, np.int16).reshape(-1, 3):
            self.lib.ora_allocate_block(self.h, int(x), int(y), int(z))

    def DeleteBlocks(self, coords):
        for x, y, z in np.asarray(coords, np.int16).reshape(-1, 3):
            self.lib.ora_delete_block(self.h, int(x), int(y), int(z))

    def Retrieve(self, pts):
        pts = np.asarray(pts, np.int16).reshape(-1, 3)
        n = len(pts)
        res = {"tsdf": np.empty(n, np.float32), "prob": np.empty(n, np.float32), "rgbw": np.empty((n, 4), np.uint8),
               "found": np.empty(n, bool)}
        o = np.zeros(6, np.float32)
        for i, (x, y, z) in enumerate(pts):
            res["found"][i] = bool(self.lib.ora_retrieve(self.h, int(x), int(y), int(z), o.ctypes.data))
            res["tsdf"][i], res["prob"][i] = o[0], o[1]
            res["rgbw"][i] = o[2:6].astype(np.uint8)
        return res

    def Assign(self, pts, tsdf=None, rgbw=None, prob=None):
        pts = np.asarray(pts, np.int16).reshape(-1, 3)
        cur = self.Retrieve(pts)
        out = np.zeros(len(pts), bool)
        for i, (x, y, z) in enumerate(pts):
            t = cur["tsdf"][i] if tsdf is None else tsdf[i]
            p = cur["prob"][i] if prob is None else prob[i]
            c = np.ascontiguousarray(cur["rgbw"][i] if rgbw is None else np.asarray(rgbw, np.uint8).reshape(-1, 4)[i])
            out[i] = bool(self.lib.ora_write_voxel(self.h, int(x), int(y), int(z), float(t), float(p), c.ctypes.data))
        return out

    def RetrieveTSDF(self, pts):
        pts = np.asarray(pts, np.float32).reshape(-1, 3)
        return np.array([self.lib.ora_trilinear(self.h, float(x), float(y), float(z)) for x, y, z in pts], np.float32)


def Hash(x, y, z, mask=(1 << 21) - 1):
    return int(load().ora_hash(x, y, z, mask))
