"""rsfusion -- B200-native voxel-hashing TSDF + semantic fusion (RA-SLAM's utils/tsdf path).

Python host mirror of the reference's operator surface, bound to the C ABI of
``include/rsfusion.h`` through ctypes:

* :class:`TSDFGrid`  -- reference ``TSDFGrid`` (utils/tsdf/voxel_tsdf.cuh:39-145): ``Integrate``,
  ``RayCast``, ``GatherValid``, ``GatherValidSemantic``, ``GatherVoxels``, ``GatherValidMesh``.
* :class:`SE3`, :class:`CameraIntrinsics`, :class:`CameraParams`, :class:`BoundingCube` -- the
  PODs of utils/cuda/{lie_group,camera}.cuh and voxel_tsdf.cuh:20-34.

The directory is called ``ra-slam_b200`` (not importable with a plain ``import`` statement):
use ``importlib.import_module("ra-slam_b200")`` or the top-level ``rsfusion`` alias module.

There is no CPU fallback: constructing a :class:`TSDFGrid` without the compiled
``librsfusion.so`` or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSF_LIBRARY") or os.path.join(_HERE, "librsfusion.so")   # RSF_LIBRARY: dev builds
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

RSF_OK, RSF_ERR_INVALID, RSF_ERR_CUDA, RSF_ERR_POOL_FULL, RSF_ERR_TABLE_FULL, RSF_ERR_CAPACITY = range(6)


class RsfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rsfusion error {code}: {msg}")
        self.code = code


class rsf_options(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_float), ("truncation", C.c_float), ("device", C.c_int),
        ("pool_blocks", C.c_int), ("table_slots_log2", C.c_int), ("max_pixels", C.c_int),
        ("shard_rank", C.c_int), ("shard_world", C.c_int), ("shard_shift", C.c_int),
        ("shard_halo", C.c_int), ("stream", C.c_void_p), ("flags", C.c_uint), ("num_classes", C.c_int),
    ]


class rsf_shard_link(C.Structure):
    _fields_ = [("planes", C.c_void_p * 2), ("flags", C.c_void_p), ("n_planes", C.c_int), ("skip_mask", C.c_uint), ("root", C.c_int),
                ("H", C.c_int), ("W", C.c_int), ("frame_H", C.c_int), ("frame_W", C.c_int), ("rx", C.c_void_p * 2),
                ("frame_flag", C.c_void_p), ("n_kids", C.c_int), ("kid_rx", (C.c_void_p * 2) * 2), ("kid_flag", C.c_void_p * 2)]


class rsf_frame_stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("active_blocks", "new_blocks", "visible_blocks", "updated_voxels", "carved_blocks", "high_water",
                 "culled_blocks", "table_keys", "table_rebuilds")]


# every symbol include/rsfusion.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F4 = C.c_void_p   # const float* arguments (K[4], pose[7], bounds[6]): passed as plain addresses, see _fptr
ABI = {
    "rsf_default_options": (None, [C.POINTER(rsf_options)]),
    "rsf_abi_version": (C.c_int, []),
    "rsf_last_error": (C.c_char_p, []),
    "rsf_create": (C.c_int, [C.POINTER(rsf_options), C.POINTER(_P)]),
    "rsf_destroy": (C.c_int, [_P]),
    "rsf_integrate": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _F4, _F4]),
    "rsf_integrate_async": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _F4, _F4]),
    "rsf_raycast_async": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, _F4, _F4, _P, _P, _P, _P, C.POINTER(C.c_int)]),
    "rsf_wait": (C.c_int, [_P, C.c_int]),
    "rsf_integrate_device": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _F4, _F4]),
    "rsf_integrate_u16": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_float, _F4, _F4]),
    "rsf_integrate_u16_async": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_float, _F4, _F4]),
    "rsf_integrate_u16_device": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_float, _F4, _F4]),
    "rsf_integrate_classes": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _F4, _F4]),
    "rsf_integrate_classes_device": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _F4, _F4]),
    "rsf_gather_valid_classes": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "rsf_resize_device": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rsf_raycast": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, _F4, _F4, _P, _P, _P, _P]),
    "rsf_raycast_device": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, _F4, _F4, _P, _P, _P, _P]),
    "rsf_raycast_internal_device": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, _F4, _F4, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "rsf_raycast_merge_device": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, _F4, _F4, _P, C.c_int]),
    "rsf_merge_finalize_device": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P]),
    "rsf_shard_set_link": (C.c_int, [_P, C.POINTER(rsf_shard_link)]),
    "rsf_shard_frame_device": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, _F4, _F4,
                                         C.c_float, C.c_int, C.c_int, _F4, _F4, _P, _P, _P, _P]),
    "rsf_shard_feed_device": (C.c_int, [_P, _P]),
    "rsf_shard_step_device": (C.c_int, [_P, C.c_float, _F4, _F4, C.c_float, _F4, _F4, _P, _P, _P, _P]),
    "rsf_stream_wait_flag_device": (C.c_int, [_P, _P, C.c_uint32]),
    "rsf_stream_wait_flag_on": (C.c_int, [C.c_int, _P, _P, C.c_uint32]),
    "rsf_synchronize": (C.c_int, [_P]),
    "rsf_num_active_blocks": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "rsf_get_stats": (C.c_int, [_P, C.POINTER(rsf_frame_stats)]),
    "rsf_gather_valid": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "rsf_gather_valid_semantic": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "rsf_gather_voxels": (C.c_int, [_P, _F4, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "rsf_gather_valid_mesh": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64), _P, C.c_int64,
                                        C.POINTER(C.c_int64), _P]),
    "rsf_export_blocks": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "rsf_hash": (C.c_uint32, [C.c_int16, C.c_int16, C.c_int16, C.c_uint32]),
    "rsf_shard_owner": (C.c_uint32, [C.c_int16, C.c_int16, C.c_int16, C.c_int, C.c_int]),
    "rsf_allocate_blocks": (C.c_int, [_P, _P, C.c_int]),
    "rsf_delete_blocks": (C.c_int, [_P, _P, C.c_int]),
    "rsf_rebuild_table": (C.c_int, [_P]),
    "rsf_set_table_policy": (C.c_int, [_P, C.c_int64]),
    "rsf_retrieve_voxels": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "rsf_assign_voxels": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "rsf_retrieve_tsdf_trilinear": (C.c_int, [_P, _P, C.c_int, _P]),
    "rsf_profile_enable": (C.c_int, [_P, C.c_int]),
    "rsf_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rsf_get_totals": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "rsf_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "rsf_host_free": (C.c_int, [_P]),
}

_lib = None


def load_library(path: str | None = None):
    """dlopen librsfusion.so and bind every symbol of include/rsfusion.h.  Raises if missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RsfError(RSF_ERR_CUDA, f"{p} not found: build it with __graft_entry__.build() "
                                     f"(make -C ra-slam_b200/csrc); there is no CPU fallback")
    lib = C.CDLL(p)
    for name, (res, args) in ABI.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.rsf_abi_version() != 2:
        raise RsfError(RSF_ERR_INVALID, "ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def _check(rc: int):
    if rc != RSF_OK:
        raise RsfError(rc, load_library().rsf_last_error().decode())


# ------------------------------------------------------------------------------------------------
# PODs of the reference interface
@dataclass(frozen=True)
class CameraIntrinsics:
    """utils/cuda/camera.cuh:13-52."""
    fx: float
    fy: float
    cx: float
    cy: float

    def as_array(self) -> np.ndarray:
        return np.array([self.fx, self.fy, self.cx, self.cy], dtype=np.float32)


@dataclass(frozen=True)
class CameraParams:
    """utils/cuda/camera.cuh:54-68."""
    intrinsics: CameraIntrinsics
    img_h: int
    img_w: int


@dataclass(frozen=True)
class BoundingCube:
    """voxel_tsdf.cuh:20-34."""
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    zmin: float
    zmax: float

    def as_array(self) -> np.ndarray:
        return np.array([self.xmin, self.xmax, self.ymin, self.ymax, self.zmin, self.zmax], dtype=np.float32)


class SE3:
    """utils/cuda/lie_group.cuh:8-45: unit quaternion (w,x,y,z) + translation, float32.

    Construction from a matrix follows Eigen 3.3's ``Quaternion(Matrix3)`` (trace-based, no
    normalisation), which is what ``SE3<float>(Matrix4f)`` does in the reference.
    """

    def __init__(self, q=(1.0, 0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0)):
        self.q = np.asarray(q, dtype=np.float32).reshape(4)
        self.t = np.asarray(t, dtype=np.float32).reshape(3)

    @staticmethod
    def Identity() -> "SE3":
        return SE3()

    @staticmethod
    def from_matrix(m) -> "SE3":
        m = np.asarray(m, dtype=np.float32)
        R, t = m[:3, :3], m[:3, 3]
        f = np.float32
        tr = f(R[0, 0] + R[1, 1] + R[2, 2])
        q = np.zeros(4, dtype=np.float32)  # w,x,y,z
        if tr > 0:
            s = np.sqrt(f(tr + f(1.0)))
            q[0] = f(0.5) * s
            s = f(0.5) / s
            q[1] = (R[2, 1] - R[1, 2]) * s
            q[2] = (R[0, 2] - R[2, 0]) * s
            q[3] = (R[1, 0] - R[0, 1]) * s
        else:
            i = 0
            if R[1, 1] > R[0, 0]:
                i = 1
            if R[2, 2] > R[i, i]:
                i = 2
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(f(R[i, i] - R[j, j] - R[k, k] + f(1.0)))
            q[1 + i] = f(0.5) * s
            s = f(0.5) / s
            q[0] = (R[k, j] - R[j, k]) * s
            q[1 + j] = (R[j, i] + R[i, j]) * s
            q[1 + k] = (R[k, i] + R[i, k]) * s
        return SE3(q, t)

    def rotation_matrix(self) -> np.ndarray:
        w, x, y, z = [float(v) for v in self.q]
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)

    def matrix(self) -> np.ndarray:
        m = np.eye(4)
        m[:3, :3] = self.rotation_matrix()
        m[:3, 3] = self.t
        return m

    def Inverse(self) -> "SE3":
        return SE3.from_matrix(np.linalg.inv(self.matrix()))

    def __mul__(self, other: "SE3") -> "SE3":
        return SE3.from_matrix(self.matrix() @ other.matrix())

    def as_array(self) -> np.ndarray:
        """{qw,qx,qy,qz,tx,ty,tz} -- the `pose` argument of the C ABI."""
        a = np.empty(7, np.float32)
        a[:4] = self.q
        a[4:] = self.t
        return a


def _fptr(a: np.ndarray):
    """Address of a contiguous float32 array (the caller keeps `a` alive across the call).  `ndarray.ctypes.data_as(POINTER)`
    costs ~5 us per argument -- a quarter of the host time of a 150 us frame -- the raw address 0.3 us."""
    return a.__array_interface__["data"][0]


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "is_cuda") and bool(x.is_cuda)


class TSDFGrid:
    """Reference ``TSDFGrid`` (utils/tsdf/voxel_tsdf.cuh:39-145) on top of the C ABI.

    ``Integrate`` / ``RayCast`` accept numpy host arrays (blocking, like the reference) or torch CUDA
    tensors (enqueued on the grid's stream; call :meth:`synchronize` or pass ``stream=`` the torch
    stream the tensors live on).
    """

    def __init__(self, voxel_size: float, truncation: float, *, device: int = -1, pool_blocks: int = 0,
                 table_slots_log2: int = 0, max_pixels: int = 0, shard_rank: int = 0, shard_world: int = 1,
                 shard_shift: int = 4, shard_halo: int = 1, stream: int | None = None, num_classes: int = 0):
        self._lib = load_library()
        o = rsf_options()
        self._lib.rsf_default_options(C.byref(o))
        o.voxel_size, o.truncation, o.device = voxel_size, truncation, device
        o.pool_blocks, o.table_slots_log2, o.max_pixels = pool_blocks, table_slots_log2, max_pixels
        o.shard_rank, o.shard_world, o.shard_shift, o.shard_halo = shard_rank, shard_world, shard_shift, shard_halo
        o.num_classes = num_classes
        self.num_classes = num_classes
        self.stream_handle = stream          # None: library-owned stream (not torch's current stream)
        if stream is not None:
            o.stream = stream or None
            o.flags = 1          # RSF_FLAG_EXTERNAL_STREAM: use it even if it is the default stream (0)
        self._h = _P()
        _check(self._lib.rsf_create(C.byref(o), C.byref(self._h)))
        self.voxel_size, self.truncation = float(voxel_size), float(truncation)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rsf_destroy(self._h)
            self._h = None
        pa = getattr(self, "_pinned_rows", None)
        if pa is not None:
            pa.free()
            self._pinned_rows = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- TSDFGrid::Integrate ------------------------------------------------------------------
    def Integrate(self, img_rgb, img_depth, img_ht, img_lt, max_depth: float, intrinsics: CameraIntrinsics,
                  cam_T_world: SE3):
        K = intrinsics.as_array()
        pose = cam_T_world.as_array()
        if _is_torch_cuda(img_depth):
            H, W = img_depth.shape
            for t, dt in ((img_rgb, "torch.uint8"), (img_depth, "torch.float32"), (img_ht, "torch.float32"),
                          (img_lt, "torch.float32")):
                assert str(t.dtype) == dt and t.is_contiguous(), "CV_8UC3 / CV_32FC1 contiguous contract"
            assert tuple(img_rgb.shape) == (H, W, 3) and img_ht.shape == img_depth.shape == img_lt.shape
            _check(self._lib.rsf_integrate_device(self._h, img_rgb.data_ptr(), img_depth.data_ptr(),
                                                  img_ht.data_ptr(), img_lt.data_ptr(), H, W, max_depth,
                                                  _fptr(K), _fptr(pose)))
            return
        # voxel_tsdf.cu:419-428: type / shape asserts
        assert img_rgb.dtype == np.uint8 and img_rgb.ndim == 3 and img_rgb.shape[2] == 3, "img_rgb must be CV_8UC3"
        for a in (img_depth, img_ht, img_lt):
            assert a.dtype == np.float32 and a.ndim == 2, "depth / ht / lt must be CV_32FC1"
            assert a.shape == img_rgb.shape[:2], "image sizes differ"
        H, W = img_depth.shape
        rgb, d, ht, lt = (np.ascontiguousarray(a) for a in (img_rgb, img_depth, img_ht, img_lt))
        _check(self._lib.rsf_integrate(self._h, rgb.ctypes.data, d.ctypes.data, ht.ctypes.data, lt.ctypes.data,
                                       H, W, max_depth, _fptr(K), _fptr(pose)))

    # -- TSDFGrid::RayCast --------------------------------------------------------------------
    def RayCast(self, max_depth: float, virtual_cam: CameraParams, cam_T_world: SE3, *, outputs=("rgba", "normal",
                "depth", "prob"), out: dict | None = None) -> dict:
        """Returns {'rgba': HxWx4 u8, 'normal': HxWx4 u8, 'depth': HxW f32, 'prob': HxW f32}.

        ``out`` may hold preallocated numpy arrays (host path) or torch CUDA tensors (device path).
        """
        H, W = virtual_cam.img_h, virtual_cam.img_w
        K = virtual_cam.intrinsics.as_array()
        pose = cam_T_world.as_array()
        if out is not None and any(_is_torch_cuda(v) for v in out.values()):
            ptr = {k: (out[k].data_ptr() if k in out else None) for k in ("rgba", "normal", "depth", "prob")}
            _check(self._lib.rsf_raycast_device(self._h, max_depth, H, W, _fptr(K), _fptr(pose), ptr["rgba"],
                                                ptr["normal"], ptr["depth"], ptr["prob"]))
            return out
        res = dict(out) if out else {}
        for k in outputs:
            if k not in res:
                res[k] = (np.empty((H, W, 4), np.uint8) if k in ("rgba", "normal") else np.empty((H, W), np.float32))
        ptr = {k: (res[k].ctypes.data if k in res else None) for k in ("rgba", "normal", "depth", "prob")}
        _check(self._lib.rsf_raycast(self._h, max_depth, H, W, _fptr(K), _fptr(pose), ptr["rgba"], ptr["normal"],
                                     ptr["depth"], ptr["prob"]))
        return res

    # -- K-class semantic extension (rsf_options.num_classes > 2) ----------------------------------
    def IntegrateClasses(self, img_rgb, img_depth, class_prob, max_depth: float, intrinsics: CameraIntrinsics, cam_T_world: SE3):
        """class_prob: K x H x W float32 per-class probabilities (numpy: blocking host path; torch CUDA: device path)."""
        H, W = img_depth.shape
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        assert tuple(class_prob.shape) == (self.num_classes, H, W)
        if _is_torch_cuda(img_depth):
            _check(self._lib.rsf_integrate_classes_device(self._h, img_rgb.data_ptr(), img_depth.data_ptr(), class_prob.data_ptr(),
                                                          H, W, max_depth, _fptr(K), _fptr(pose)))
        else:
            img_rgb, img_depth, class_prob = (np.ascontiguousarray(a) for a in (img_rgb, img_depth, class_prob))
            assert img_rgb.dtype == np.uint8 and img_depth.dtype == np.float32 and class_prob.dtype == np.float32
            _check(self._lib.rsf_integrate_classes(self._h, img_rgb.ctypes.data, img_depth.ctypes.data, class_prob.ctypes.data,
                                                   H, W, max_depth, _fptr(K), _fptr(pose)))

    def GatherValidClasses(self, pinned: bool = False) -> np.ndarray:
        """N x (4 + K) float32 rows {x, y, z, tsdf, p_0 .. p_{K-1}}."""
        return self._gather(self._lib.rsf_gather_valid_classes, 4 + self.num_classes, pinned=pinned)

    def RayCastInternal(self, max_depth: float, virtual_cam: CameraParams, cam_T_world: SE3):
        """TSDFGrid::RayCast as the reference's callers use it (voxel_tsdf.cu:885-902): render into the grid's own device
        images (the reference's img_tsdf_rgba_ / img_tsdf_normal_, handed to GLImage8UC4::LoadCuda) -- no read-back.
        Returns (rgba device address, normal device address, cudaStream_t); asynchronous, `synchronize()` to wait."""
        K, pose = virtual_cam.intrinsics.as_array(), cam_T_world.as_array()
        a, b, st = _P(), _P(), _P()
        _check(self._lib.rsf_raycast_internal_device(self._h, max_depth, virtual_cam.img_h, virtual_cam.img_w, _fptr(K), _fptr(pose),
                                                     C.byref(a), C.byref(b), C.byref(st)))
        return a.value, b.value, st.value

    def RayCastMerge(self, max_depth: float, virtual_cam: CameraParams, cam_T_world: SE3, merge_ptr: int, n_planes: int = 3):
        """Sharded grids: raycast with the min-depth merge fused in (rsf_raycast_merge_device).  `merge_ptr` is the
        device address of the presenting rank's planes (peer memory on the other ranks)."""
        K, pose = virtual_cam.intrinsics.as_array(), cam_T_world.as_array()
        _check(self._lib.rsf_raycast_merge_device(self._h, max_depth, virtual_cam.img_h, virtual_cam.img_w, _fptr(K),
                                                  _fptr(pose), merge_ptr, n_planes))

    def MergeFinalize(self, H: int, W: int, merge_ptr: int, n_planes: int, out: dict):
        """Presenting rank: unpack the merged planes into the torch CUDA tensors of `out` and re-arm the planes."""
        ptr = {k: (out[k].data_ptr() if k in out else None) for k in ("rgba", "normal", "depth", "prob")}
        _check(self._lib.rsf_merge_finalize_device(self._h, H, W, merge_ptr, n_planes, ptr["rgba"], ptr["normal"],
                                                   ptr["depth"], ptr["prob"]))
        return out

    def ShardSetLink(self, planes0: int, planes1: int, flags: int, n_planes: int, root: int, H: int, W: int, skip_mask: int = 0,
                     transport: dict | None = None):
        """rsf_shard_set_link: device addresses (in THIS process) of the presenting rank's merge planes / arrival flags.
        transport = {frame_H, frame_W, rx: [p0, p1], frame_flag: p, kids: [{rx: [p0, p1], flag: p}, ...]} enables the
        in-library frame tree (ShardFeed / ShardStep)."""
        ln = rsf_shard_link()
        ln.planes[0], ln.planes[1], ln.flags = planes0, planes1, flags
        ln.n_planes, ln.root, ln.H, ln.W, ln.skip_mask = n_planes, root, H, W, skip_mask
        if transport:
            ln.frame_H, ln.frame_W = transport["frame_H"], transport["frame_W"]
            ln.rx[0], ln.rx[1], ln.frame_flag = transport["rx"][0], transport["rx"][1], transport["frame_flag"]
            ln.n_kids = len(transport["kids"])
            for c, kid in enumerate(transport["kids"]):
                ln.kid_rx[c][0], ln.kid_rx[c][1], ln.kid_flag[c] = kid["rx"][0], kid["rx"][1], kid["flag"]
        _check(self._lib.rsf_shard_set_link(self._h, C.byref(ln)))

    def ShardFeed(self, packed_frame_ptr: int | None):
        _check(self._lib.rsf_shard_feed_device(self._h, packed_frame_ptr))

    def ShardStep(self, max_depth: float, K: np.ndarray, pose: np.ndarray, raycast_max_depth: float, rc_K: np.ndarray,
                  rc_pose: np.ndarray, out_ptrs=(None, None, None, None)):
        """rsf_shard_step_device with pre-built float32 arrays (K[4], pose[7]) and raw output pointers (rgba, normal, depth, prob)."""
        _check(self._lib.rsf_shard_step_device(self._h, max_depth, _fptr(K), _fptr(pose), raycast_max_depth, _fptr(rc_K), _fptr(rc_pose),
                                               *out_ptrs))

    def ShardFrame(self, img_rgb, img_depth, img_ht, img_lt, max_depth: float, intrinsics: CameraIntrinsics, cam_T_world: SE3,
                   raycast_max_depth: float, virtual_cam: CameraParams, rc_cam_T_world: SE3, out: dict | None = None):
        """rsf_shard_frame_device: integrate + raycast-merge + arrival flag (+ unpack on the presenting rank) in one
        call; torch CUDA tensors in, `out` (torch CUDA tensors, presenting rank only) filled asynchronously."""
        H, W = img_depth.shape
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        rK, rpose = virtual_cam.intrinsics.as_array(), rc_cam_T_world.as_array()
        out = out or {}
        ptr = {k: (out[k].data_ptr() if k in out else None) for k in ("rgba", "normal", "depth", "prob")}
        _check(self._lib.rsf_shard_frame_device(self._h, img_rgb.data_ptr(), img_depth.data_ptr(), img_ht.data_ptr(),
                                                img_lt.data_ptr(), H, W, max_depth, _fptr(K), _fptr(pose), raycast_max_depth,
                                                virtual_cam.img_h, virtual_cam.img_w, _fptr(rK), _fptr(rpose),
                                                ptr["rgba"], ptr["normal"], ptr["depth"], ptr["prob"]))
        return out

    def StreamWaitFlag(self, flag_ptr: int, value: int):
        """rsf_stream_wait_flag_device: the grid's stream waits (on the device) until *flag_ptr >= value."""
        _check(self._lib.rsf_stream_wait_flag_device(self._h, flag_ptr, value))

    def synchronize(self):
        _check(self._lib.rsf_synchronize(self._h))

    # -- pipelined host path (rsf_integrate_async / rsf_raycast_async / rsf_wait) ----------------
    def IntegrateAsync(self, img_rgb, img_depth, img_ht, img_lt, max_depth: float, intrinsics: CameraIntrinsics,
                       cam_T_world: SE3):
        """Enqueue only.  The (ideally pinned) numpy buffers must stay alive and unchanged until
        `synchronize()` / `wait()` returns."""
        H, W = img_depth.shape
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        _check(self._lib.rsf_integrate_async(self._h, img_rgb.ctypes.data, img_depth.ctypes.data, img_ht.ctypes.data,
                                             img_lt.ctypes.data, H, W, max_depth, _fptr(K), _fptr(pose)))

    def IntegrateU16(self, img_rgb, depth_u16, ht_u16, lt_u16, depth_factor: float, max_depth: float,
                     intrinsics: CameraIntrinsics, cam_T_world: SE3, *, asynchronous: bool = False):
        """Integrate from raw 16-bit depth / ht / lt images (rsf_integrate_u16*): numpy arrays (host; blocking or
        `asynchronous`) or torch CUDA tensors (device).  ht_u16 = lt_u16 = None means "no segmentation"."""
        H, W = depth_u16.shape
        K, pose = intrinsics.as_array(), cam_T_world.as_array()
        if _is_torch_cuda(depth_u16):
            ptrs = [img_rgb.data_ptr(), depth_u16.data_ptr(), ht_u16.data_ptr() if ht_u16 is not None else None,
                    lt_u16.data_ptr() if lt_u16 is not None else None]
            fn = self._lib.rsf_integrate_u16_device
        else:
            assert img_rgb.dtype == np.uint8 and depth_u16.dtype == np.uint16
            assert all(a.flags["C_CONTIGUOUS"] for a in (img_rgb, depth_u16))
            ptrs = [img_rgb.ctypes.data, depth_u16.ctypes.data, ht_u16.ctypes.data if ht_u16 is not None else None,
                    lt_u16.ctypes.data if lt_u16 is not None else None]
            fn = self._lib.rsf_integrate_u16_async if asynchronous else self._lib.rsf_integrate_u16
        _check(fn(self._h, *ptrs, H, W, depth_factor, max_depth, _fptr(K), _fptr(pose)))

    def RayCastAsync(self, max_depth: float, virtual_cam: CameraParams, cam_T_world: SE3, out: dict) -> int:
        """Enqueue raycast + read-back into the numpy arrays of `out`; returns a ticket for `wait`."""
        K, pose = virtual_cam.intrinsics.as_array(), cam_T_world.as_array()
        ptr = {k: (out[k].ctypes.data if k in out else None) for k in ("rgba", "normal", "depth", "prob")}
        t = C.c_int(0)
        _check(self._lib.rsf_raycast_async(self._h, max_depth, virtual_cam.img_h, virtual_cam.img_w, _fptr(K),
                                           _fptr(pose), ptr["rgba"], ptr["normal"], ptr["depth"], ptr["prob"],
                                           C.byref(t)))
        return t.value

    def wait(self, ticket: int):
        _check(self._lib.rsf_wait(self._h, ticket))

    # -- exports --------------------------------------------------------------------------------
    def _gather(self, fn, ncol, *pre, pinned: bool = False):
        n = C.c_int64(0)
        _check(fn(self._h, *pre, None, 0, C.byref(n)))
        if pinned:
            # grid-owned pinned buffer (grown on demand, reused): the DMA lands in it directly, no host copy, no page
            # faults on a fresh destination.  The returned array is a VIEW, valid until the next pinned gather / close().
            need = n.value * ncol
            pa = getattr(self, "_pinned_rows", None)
            if pa is None or pa.array.size < need:
                if pa is not None:
                    pa.free()
                pa = self._pinned_rows = PinnedArray((max(need + need // 4, 1 << 20),), np.float32)
            rows = pa.array[:need].reshape(n.value, ncol)
        else:
            rows = np.empty((n.value, ncol), np.float32)
        if n.value:
            _check(fn(self._h, *pre, rows.ctypes.data, n.value, C.byref(n)))
        return rows

    def GatherValid(self, pinned: bool = False) -> np.ndarray:
        """N x 4 float32 rows {x,y,z,tsdf} (VoxelSpatialTSDF)."""
        return self._gather(self._lib.rsf_gather_valid, 4, pinned=pinned)

    def GatherValidSemantic(self, pinned: bool = False) -> np.ndarray:
        """N x 5 float32 rows {x,y,z,tsdf,prob} (VoxelSpatialTSDFSEGM)."""
        return self._gather(self._lib.rsf_gather_valid_semantic, 5, pinned=pinned)

    def GatherVoxels(self, volumn: BoundingCube) -> np.ndarray:
        b = volumn.as_array()
        return self._gather(self._lib.rsf_gather_voxels, 4, _fptr(b))

    def GatherValidMesh(self):
        """(vertices Nv x 3 f32, triangles Nt x 3 i32, vertex_prob Nv f32)."""
        nv, nt = C.c_int64(0), C.c_int64(0)
        _check(self._lib.rsf_gather_valid_mesh(self._h, None, 0, C.byref(nv), None, 0, C.byref(nt), None))
        v = np.empty((nv.value, 3), np.float32)
        t = np.empty((nt.value, 3), np.int32)
        p = np.empty((nv.value,), np.float32)
        if nv.value or nt.value:
            _check(self._lib.rsf_gather_valid_mesh(self._h, v.ctypes.data, nv.value, C.byref(nv), t.ctypes.data,
                                                   nt.value, C.byref(nt), p.ctypes.data))
        return v, t, p

    def NumActiveBlock(self) -> int:
        n = C.c_int64(0)
        _check(self._lib.rsf_num_active_blocks(self._h, C.byref(n)))
        return n.value

    def stats(self) -> dict:
        s = rsf_frame_stats()
        _check(self._lib.rsf_get_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in rsf_frame_stats._fields_}

    PROF_KERNELS = ("allocate", "select", "integrate", "raycast", "merge")

    def profile_enable(self, on: bool = True):
        _check(self._lib.rsf_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        """{kernel: (total_ms, launches)} since the last read."""
        ms = (C.c_double * len(self.PROF_KERNELS))()
        n = (C.c_int64 * len(self.PROF_KERNELS))()
        _check(self._lib.rsf_profile_read(self._h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROF_KERNELS)}

    def totals(self) -> dict:
        t = (C.c_int64 * 5)()
        _check(self._lib.rsf_get_totals(self._h, t))
        return dict(zip(("frames", "new_blocks", "visible_blocks", "updated_voxels", "carved_blocks"), list(t)))

    def ExportBlocks(self, owned_only: bool = False, sort: bool = True):
        """(coords n x 3 i16, tsdf n x 512 f32, rgbw n x 512 x 4 u8, prob n x 512 f32); sorted by key."""
        n = C.c_int64(0)
        _check(self._lib.rsf_export_blocks(self._h, int(owned_only), None, None, None, None, 0, C.byref(n)))
        nb = n.value
        coords = np.empty((nb, 3), np.int16)
        tsdf = np.empty((nb, 512), np.float32)
        rgbw = np.empty((nb, 512, 4), np.uint8)
        prob = np.empty((nb, 512), np.float32)
        if nb:
            _check(self._lib.rsf_export_blocks(self._h, int(owned_only), coords.ctypes.data, tsdf.ctypes.data,
                                               rgbw.ctypes.data, prob.ctypes.data, nb, C.byref(n)))
        if sort and nb:
            order = np.argsort(block_keys(coords), kind="stable")
            coords, tsdf, rgbw, prob = coords[order], tsdf[order], rgbw[order], prob[order]
        return coords, tsdf, rgbw, prob

    def ExportBlockCoords(self, owned_only: bool = False) -> np.ndarray:
        """n x 3 i16 block coordinates of the alive blocks (no voxel data)."""
        n = C.c_int64(0)
        _check(self._lib.rsf_export_blocks(self._h, int(owned_only), None, None, None, None, 0, C.byref(n)))
        coords = np.empty((n.value, 3), np.int16)
        if n.value:
            _check(self._lib.rsf_export_blocks(self._h, int(owned_only), coords.ctypes.data, None, None, None, n.value, C.byref(n)))
        return coords

    # -- table / pool level hooks (utils/tests/voxel_hash_test.cu) -------------------------------
    def AllocateBlocks(self, block_coords):
        c = np.ascontiguousarray(block_coords, dtype=np.int16).reshape(-1, 3)
        _check(self._lib.rsf_allocate_blocks(self._h, c.ctypes.data, len(c)))

    def DeleteBlocks(self, block_coords):
        c = np.ascontiguousarray(block_coords, dtype=np.int16).reshape(-1, 3)
        _check(self._lib.rsf_delete_blocks(self._h, c.ctypes.data, len(c)))

    def Resize(self, src, size):
        """cv::resize(src, Size(W, H)) (INTER_LINEAR) of a torch CUDA tensor -- uint8 HxWx3 / HxW, uint16 or int16 HxW bits,
        float32 HxW -- on the grid's stream (rsf_resize_device).  size = (H, W).  Returns a new tensor."""
        import torch
        assert _is_torch_cuda(src) and src.is_contiguous()
        ch = src.shape[2] if src.dim() == 3 else 1
        dst = torch.empty((size[0], size[1]) + ((ch,) if src.dim() == 3 else ()), dtype=src.dtype, device=src.device)
        _check(self._lib.rsf_resize_device(self._h, src.data_ptr(), src.shape[0], src.shape[1], dst.data_ptr(), size[0], size[1],
                                           ch, src.element_size()))
        return dst

    def RebuildTable(self):
        """Re-insert the alive blocks into an empty table (drops the ghost keys of carved blocks); the library does this
        by itself when the ghost load passes its threshold."""
        _check(self._lib.rsf_rebuild_table(self._h))

    def SetTablePolicy(self, max_ghosts: int):
        _check(self._lib.rsf_set_table_policy(self._h, int(max_ghosts)))

    def Retrieve(self, voxel_coords):
        c = np.ascontiguousarray(voxel_coords, dtype=np.int16).reshape(-1, 3)
        n = len(c)
        tsdf, prob = np.empty(n, np.float32), np.empty(n, np.float32)
        rgbw, found = np.empty((n, 4), np.uint8), np.empty(n, np.uint8)
        _check(self._lib.rsf_retrieve_voxels(self._h, c.ctypes.data, n, tsdf.ctypes.data, rgbw.ctypes.data,
                                             prob.ctypes.data, found.ctypes.data))
        return {"tsdf": tsdf, "rgbw": rgbw, "prob": prob, "found": found.astype(bool)}

    def Assign(self, voxel_coords, tsdf=None, rgbw=None, prob=None):
        c = np.ascontiguousarray(voxel_coords, dtype=np.int16).reshape(-1, 3)
        n = len(c)
        t = None if tsdf is None else np.ascontiguousarray(tsdf, np.float32).reshape(n)
        r = None if rgbw is None else np.ascontiguousarray(rgbw, np.uint8).reshape(n, 4)
        p = None if prob is None else np.ascontiguousarray(prob, np.float32).reshape(n)
        written = np.empty(n, np.uint8)
        _check(self._lib.rsf_assign_voxels(self._h, c.ctypes.data, n, None if t is None else t.ctypes.data,
                                           None if r is None else r.ctypes.data,
                                           None if p is None else p.ctypes.data, written.ctypes.data))
        return written.astype(bool)

    def RetrieveTSDF(self, points):
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(p), np.float32)
        _check(self._lib.rsf_retrieve_tsdf_trilinear(self._h, p.ctypes.data, len(p), out.ctypes.data))
        return out


def stream_wait_flag(device: int, stream: int, flag_ptr: int, value: int):
    """rsf_stream_wait_flag_on: cudaStream_t `stream` of `device` waits (on the device) until *flag_ptr >= value."""
    _check(load_library().rsf_stream_wait_flag_on(device, stream or None, flag_ptr, value))


def block_keys(coords: np.ndarray) -> np.ndarray:
    """Packed 48-bit block key x | y<<16 | z<<32 (unsigned 16-bit fields), as the table stores it."""
    c = np.asarray(coords).astype(np.int16).view(np.uint16).astype(np.uint64).reshape(-1, 3)
    return c[:, 0] | (c[:, 1] << np.uint64(16)) | (c[:, 2] << np.uint64(32))


def Hash(x: int, y: int, z: int, mask: int = (1 << 21) - 1) -> int:
    """utils/tsdf/voxel_hash.cu:19-23 (host-callable; does not need a GPU)."""
    return int(load_library().rsf_hash(x, y, z, mask))


def shard_owner(x: int, y: int, z: int, shift: int, world: int) -> int:
    return int(load_library().rsf_shard_owner(x, y, z, shift, world))


class PinnedArray:
    """numpy view of page-locked host memory from rsf_host_alloc (true async DMA for Integrate)."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = _P()
        _check(self._lib.rsf_host_alloc(C.byref(self._p), max(self.nbytes, 1)))
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            self._lib.rsf_host_free(self._p)
            self._p = None
