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
