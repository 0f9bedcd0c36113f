"""On-disk formats either side of the TSDF path, host side only -- no kernels here.

* the folder dataset the reference's offline tools read (utils/offline_data_provider/folder_reader.cc:9-105,
  examples/tsdf/offline.cc:45-83): ``<id>_rgb.png`` (8-bit BGR on disk), ``<id>_depth.png`` (uint16, metres =
  raw / depthmap_factor), optional ``<id>_ht.png`` / ``<id>_no_ht.png`` (uint16, probability = raw / 65535),
  ``trajectory.txt`` (``id`` + 12 floats = row-major 3x4 cam_T_world per line) and ``camera_config.yaml``
  (``Camera.fx/fy/cx/cy``, ``depthmap_factor``, optional row-major 4x4 ``Extrinsics`` applied as
  ``extrinsics * cam_T_world``);
* the dumps the evaluation scripts read back: the ``N x 5`` float32 file of ``GatherValidSemantic`` rows
  (examples/scannet_evaluation/eval_one.cc:82-86, python_utils/scannet_eval/scanneteval.py:149-177) and the three mesh
  files of ``TSDFSystem::DownloadAllMesh`` (modules/tsdf_module.cc:66-86, python_utils/scannet_eval/mesh_processor.py:10-15).

The raw uint16 images go straight to ``TSDFGrid.IntegrateU16`` (conversion on the GPU), or through ``frame()`` which
converts like ``cv::Mat::convertTo`` for the float entry point.
"""
from __future__ import annotations

import os

import numpy as np

from . import CameraIntrinsics, SE3


# ---- trajectory.txt -------------------------------------------------------------------------------------------
def read_trajectory(path: str, extrinsics: SE3 | None = None):
    """[(id, cam_T_world)] -- folder_reader.cc:86-105: each line ``id m00 .. m23`` (3x4 row-major)."""
    out = []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if len(tok) < 13:
                continue
            m = np.eye(4)
            m[:3, :4] = np.array(tok[1:13], dtype=np.float64).reshape(3, 4)
            pose = SE3.from_matrix(m)
            out.append((int(tok[0]), extrinsics * pose if extrinsics is not None else pose))
    return out


def write_trajectory(path: str, entries):
    with open(path, "w") as f:
        for fid, pose in entries:
            m = pose.matrix()[:3, :4].reshape(-1)
            f.write(str(int(fid)) + " " + " ".join(f"{v:.9g}" for v in m) + "\n")


# ---- dumps ----------------------------------------------------------------------------------------------------
def write_semantic_bin(path: str, rows: np.ndarray):
    """GatherValidSemantic rows (x, y, z, tsdf, prob) as eval_one.cc writes them: raw float32, 20 bytes per voxel."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    assert rows.ndim == 2 and rows.shape[1] == 5
    rows.tofile(path)


def read_semantic_bin(path: str) -> np.ndarray:
    return np.fromfile(path, dtype=np.float32).reshape(-1, 5)


def write_mesh_bins(directory: str, vertices: np.ndarray, triangles: np.ndarray, vertex_prob: np.ndarray):
    """mesh_vertices.bin (f32 x 3), mesh_indices.bin (i32 x 3), mesh_vertices_prob.bin (f32): TSDFSystem::DownloadAllMesh."""
    os.makedirs(directory, exist_ok=True)
    np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3).tofile(os.path.join(directory, "mesh_vertices.bin"))
    np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3).tofile(os.path.join(directory, "mesh_indices.bin"))
    np.ascontiguousarray(vertex_prob, dtype=np.float32).reshape(-1).tofile(os.path.join(directory, "mesh_vertices_prob.bin"))


def read_mesh_bins(directory: str):
    v = np.fromfile(os.path.join(directory, "mesh_vertices.bin"), dtype=np.float32).reshape(-1, 3)
    t = np.fromfile(os.path.join(directory, "mesh_indices.bin"), dtype=np.int32).reshape(-1, 3)
    p = np.fromfile(os.path.join(directory, "mesh_vertices_prob.bin"), dtype=np.float32).reshape(-1)
    return v, t, p


# ---- folder dataset -------------------------------------------------------------------------------------------
class FolderDataset:
    """Reader (and writer, for synthetic data) of the reference's folder dataset."""

    def __init__(self, folder: str):
        import yaml
        self.folder = folder
        with open(os.path.join(folder, "camera_config.yaml")) as f:
            cfg = yaml.safe_load(f)
        self.intrinsics = CameraIntrinsics(float(cfg["Camera.fx"]), float(cfg["Camera.fy"]), float(cfg["Camera.cx"]),
                                           float(cfg["Camera.cy"]))
        self.depth_factor = float(cfg["depthmap_factor"])
        ext = cfg.get("Extrinsics") or []
        self.extrinsics = SE3.from_matrix(np.array(ext, dtype=np.float64).reshape(4, 4)) if len(ext) == 16 else None
        self.entries = read_trajectory(os.path.join(folder, "trajectory.txt"), self.extrinsics)

    def __len__(self):
        return len(self.entries)

    def pose(self, i: int) -> SE3:
        return self.entries[i][1]

    def raw(self, i: int):
        """(rgb u8 HxWx3 in RGB order, depth u16, ht u16 | None, lt u16 | None) -- what IntegrateU16 takes."""
        import cv2
        fid = self.entries[i][0]
        bgr = cv2.imread(os.path.join(self.folder, f"{fid}_rgb.png"))
        rgb = np.ascontiguousarray(bgr[..., ::-1])                       # folder_reader.cc:70 COLOR_BGR2RGB
        depth = cv2.imread(os.path.join(self.folder, f"{fid}_depth.png"), cv2.IMREAD_UNCHANGED)
        ht = cv2.imread(os.path.join(self.folder, f"{fid}_ht.png"), cv2.IMREAD_UNCHANGED)
        lt = cv2.imread(os.path.join(self.folder, f"{fid}_no_ht.png"), cv2.IMREAD_UNCHANGED)
        assert depth is not None and depth.dtype == np.uint16 and depth.shape == rgb.shape[:2]
        return rgb, depth, ht, lt

    def frame(self, i: int):
        """(rgb, depth f32 [m], ht f32, lt f32) converted like offline.cc:62-76 (convertTo multiplies in float)."""
        rgb, depth, ht, lt = self.raw(i)
        d = depth.astype(np.float32) * np.float32(1.0 / self.depth_factor)
        if ht is not None:
            return rgb, d, ht.astype(np.float32) * np.float32(1.0 / 65535), lt.astype(np.float32) * np.float32(1.0 / 65535)
        return rgb, d, np.zeros_like(d), np.ones_like(d)

    @staticmethod
    def write(folder: str, intrinsics: CameraIntrinsics, depth_factor: float, frames, poses, ids=None):
        """frames: iterable of (rgb u8, depth f32 [m], ht f32, lt f32); quantises like a real recorder would."""
        import cv2
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, "camera_config.yaml"), "w") as f:
            f.write(f"Camera.fx: {intrinsics.fx}\nCamera.fy: {intrinsics.fy}\nCamera.cx: {intrinsics.cx}\n"
                    f"Camera.cy: {intrinsics.cy}\ndepthmap_factor: {depth_factor}\n")
        entries = []
        for n, ((rgb, depth, ht, lt), pose) in enumerate(zip(frames, poses)):
            fid = ids[n] if ids is not None else n
            cv2.imwrite(os.path.join(folder, f"{fid}_rgb.png"), np.ascontiguousarray(rgb[..., ::-1]))
            cv2.imwrite(os.path.join(folder, f"{fid}_depth.png"),
                        np.clip(np.rint(depth.astype(np.float64) * depth_factor), 0, 65535).astype(np.uint16))
            cv2.imwrite(os.path.join(folder, f"{fid}_ht.png"), np.clip(np.rint(ht.astype(np.float64) * 65535), 0, 65535).astype(np.uint16))
            cv2.imwrite(os.path.join(folder, f"{fid}_no_ht.png"), np.clip(np.rint(lt.astype(np.float64) * 65535), 0, 65535).astype(np.uint16))
            entries.append((fid, pose))
        write_trajectory(os.path.join(folder, "trajectory.txt"), entries)


# ---- ScanNet .sens ----------------------------------------------------------------------------------------------
class SensReader:
    """ScanNet ``.sens`` container, as utils/offline_data_provider/scannet_sens_reader.cc:5-101 reads it through
    third_party/scannet/sensorData.hpp:491-531 (version 4): header = version u32, name (u64 length + bytes), colour and
    depth calibration (intrinsic + extrinsic, 4 x 16 f32 row-major), colour / depth compression i32 (colour 2 = JPEG;
    depth 0 = raw u16, 1 = zlib u16), colour W H and depth W H (u32), depth shift (f32), frame count u64; per frame
    (RGBDFrame.h:286-297) camera_to_world 16 f32 row-major, two u64 time stamps, u64 colour bytes, u64 depth bytes, the
    two payloads.  Frames are indexed on open and decoded on demand.

    Like the reference reader: intrinsics / depth factor come from the DEPTH calibration, the colour image is resized to
    the depth size with cv::resize (:78), and the pose is the inverse of camera_to_world (:91-99).  The depth payload is
    inflated on the CPU (zlib); the uint16 image goes to ``TSDFGrid.IntegrateU16`` unchanged.  JPEG decoding stays on the
    CPU (cv2; the reference uses stb_image, so colours may differ by the decoders' IDCT rounding)."""

    COLOR_RAW, COLOR_PNG, COLOR_JPEG = 0, 1, 2
    DEPTH_RAW, DEPTH_ZLIB, DEPTH_OCCI = 0, 1, 2

    def __init__(self, path: str):
        import struct
        self.path = path
        self._f = open(path, "rb")
        f = self._f
        (self.version,) = struct.unpack("<I", f.read(4))
        if self.version != 4:
            raise ValueError(f"{path}: .sens version {self.version}, expected 4")          # sensorData.hpp:125-126
        (n,) = struct.unpack("<Q", f.read(8))
        self.sensor_name = f.read(n).decode("latin-1")
        mats = np.frombuffer(f.read(4 * 64), dtype="<f4").reshape(4, 4, 4)
        self.color_intrinsic, self.color_extrinsic, self.depth_intrinsic, self.depth_extrinsic = (m.copy() for m in mats)
        self.color_compression, self.depth_compression = struct.unpack("<ii", f.read(8))
        self.color_width, self.color_height, self.depth_width, self.depth_height = struct.unpack("<IIII", f.read(16))
        (self.depth_shift,) = struct.unpack("<f", f.read(4))
        (count,) = struct.unpack("<Q", f.read(8))
        self._index = []
        for _ in range(count):
            c2w = np.frombuffer(f.read(64), dtype="<f4").reshape(4, 4).copy()
            ts_c, ts_d, nc, nd = struct.unpack("<QQQQ", f.read(32))
            off = f.tell()
            self._index.append((c2w, ts_c, ts_d, off, nc, off + nc, nd))
            f.seek(nc + nd, os.SEEK_CUR)

    def close(self):
        self._f.close()

    def __len__(self):
        return len(self._index)

    # scannet_sens_reader.cc:9-15, :38, :101-105
    @property
    def intrinsics(self) -> CameraIntrinsics:
        m = self.depth_intrinsic
        return CameraIntrinsics(float(m[0, 0]), float(m[1, 1]), float(m[0, 2]), float(m[1, 2]))

    @property
    def depth_factor(self) -> float:
        return float(self.depth_shift)

    @property
    def size(self):
        return self.depth_height, self.depth_width

    def pose(self, i: int) -> SE3:
        """cam_T_world = SE3(camera_to_world).Inverse() (:91-99)."""
        return SE3.from_matrix(self._index[i][0]).Inverse()

    def depth_raw(self, i: int) -> np.ndarray:
        """uint16 depth image (:40-54), raw units (metres = raw / depth_factor)."""
        import zlib
        _, _, _, _, _, off, n = self._index[i]
        self._f.seek(off)
        buf = self._f.read(n)
        if self.depth_compression == self.DEPTH_ZLIB:
            buf = zlib.decompress(buf)
        elif self.depth_compression != self.DEPTH_RAW:
            raise NotImplementedError("OCCI depth compression (not used by ScanNet releases)")
        return np.frombuffer(buf, dtype="<u2").reshape(self.depth_height, self.depth_width).copy()

    def color_full(self, i: int) -> np.ndarray:
        """RGB uint8 image at the colour sensor's resolution (1296 x 968 in ScanNet)."""
        import cv2
        _, _, _, off, n, _, _ = self._index[i]
        self._f.seek(off)
        buf = self._f.read(n)
        if self.color_compression == self.COLOR_RAW:
            return np.frombuffer(buf, dtype=np.uint8).reshape(self.color_height, self.color_width, 3).copy()
        bgr = cv2.imdecode(np.frombuffer(buf, dtype=np.uint8), cv2.IMREAD_COLOR)
        return np.ascontiguousarray(bgr[..., ::-1])

    def color(self, i: int, resize=None) -> np.ndarray:
        """Colour image resized to the depth size like :78 (`resize` = callable(img, (H, W)), default cv2.resize; pass
        a GPU resize to keep that step off the CPU)."""
        img = self.color_full(i)
        if img.shape[:2] == self.size:
            return img
        if resize is not None:
            return resize(img, self.size)
        import cv2
        return cv2.resize(img, (self.depth_width, self.depth_height))

    def raw(self, i: int):
        """(rgb u8 HxWx3, depth u16, None, None) -- what IntegrateU16 takes (ScanNet ships no ht / lt images)."""
        return self.color(i), self.depth_raw(i), None, None


def write_sens(path: str, intrinsics: CameraIntrinsics, depth_shift: float, frames, cam_T_world, *, color_size=None,
               jpeg_quality: int = 95, raw_color: bool = False, sensor_name: str = "synthetic"):
    """Write a version-4 ``.sens`` file (sensorData.hpp:296-330 layout) from (rgb u8, depth f32 [m]) frames: zlib u16 depth,
    JPEG (or raw) colour, optionally up-sampled to `color_size` = (H, W) first, as ScanNet's colour stream is larger than
    its depth stream.  For tests and synthetic datasets."""
    import struct
    import zlib
    import cv2
    frames = list(frames)
    dH, dW = frames[0][1].shape
    cH, cW = color_size if color_size else (dH, dW)
    K = np.eye(4, dtype="<f4")
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy
    Kc = K.copy()
    Kc[0] *= cW / dW
    Kc[1] *= cH / dH
    eye = np.eye(4, dtype="<f4")
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 4))
        f.write(struct.pack("<Q", len(sensor_name)) + sensor_name.encode("latin-1"))
        for m in (Kc, eye, K, eye):
            f.write(np.ascontiguousarray(m, dtype="<f4").tobytes())
        f.write(struct.pack("<ii", SensReader.COLOR_RAW if raw_color else SensReader.COLOR_JPEG, SensReader.DEPTH_ZLIB))
        f.write(struct.pack("<IIII", cW, cH, dW, dH))
        f.write(struct.pack("<f", float(depth_shift)))
        f.write(struct.pack("<Q", len(frames)))
        for i, ((rgb, depth), pose) in enumerate(zip(frames, cam_T_world)):
            c2w = np.linalg.inv(pose.matrix()).astype("<f4")
            if (cH, cW) != rgb.shape[:2]:
                rgb = cv2.resize(rgb, (cW, cH))
            if raw_color:
                cbuf = np.ascontiguousarray(rgb).tobytes()
            else:
                ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, jpeg_quality])
                assert ok
                cbuf = enc.tobytes()
            d16 = np.clip(np.rint(depth.astype(np.float64) * depth_shift), 0, 65535).astype("<u2")
            dbuf = zlib.compress(d16.tobytes())
            f.write(c2w.tobytes())
            f.write(struct.pack("<QQQQ", i, i, len(cbuf), len(dbuf)))
            f.write(cbuf)
            f.write(dbuf)
        f.write(struct.pack("<Q", 0))          # no IMU frames
