"""Synthetic RGB-D + semantic sequences with closed-form depth.

Scenes are axis-aligned: a box room (camera inside) plus solid boxes; several rooms joined by
door openings make the "building".  Depth is exact z-depth along the optical axis (what the
reference's depth images hold, voxel_tsdf.cu:146), RGB is a 0.25 m checker on world coordinates,
the high-touch probability is 0.9 on "touch" surfaces (solid boxes and one wall) and 0.1 elsewhere,
lt = 1 - ht.  Everything is deterministic given the seed; numpy only (bench / tests / smoke).

Camera convention: x right, y down, z forward; world z up.  ``cam_T_world`` maps world to camera.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import CameraIntrinsics, SE3

# intrinsics of the reference's configs
TUM_FR1 = CameraIntrinsics(517.306408, 516.469215, 318.643040, 255.313989)     # configs/TUM_RGBD_rgbd_1.yaml:11-14
SCANNET = CameraIntrinsics(571.623718, 571.623718, 319.5, 239.5)               # configs/scannet_scene0.yaml:12-15
L515 = CameraIntrinsics(456.8617, 456.77127, 322.1042, 187.79485)              # configs/zed_native_l515.yaml:30-34


def scaled(K: CameraIntrinsics, sx: float, sy: float) -> CameraIntrinsics:
    return CameraIntrinsics(K.fx * sx, K.fy * sy, (K.cx + 0.5) * sx - 0.5, (K.cy + 0.5) * sy - 0.5)


@dataclass
class Scene:
    room_min: np.ndarray                      # (3,) the enclosing box the camera lives in
    room_max: np.ndarray
    boxes: list = field(default_factory=list)  # [(min(3), max(3))] solid boxes
    touch_wall_axis: int = 0                   # the wall x = room_max[0] is "high touch"


def room_scene(size=(6.0, 4.0, 3.0), n_boxes: int = 3, seed: int = 0) -> Scene:
    rng = np.random.default_rng(seed)
    lo = np.array([0.013, 0.017, 0.011])        # off-lattice so surfaces do not sit on voxel ties
    hi = lo + np.asarray(size, dtype=np.float64)
    boxes = []
    for _ in range(n_boxes):
        s = rng.uniform([0.4, 0.4, 0.4], [1.0, 1.0, 1.2])
        c = rng.uniform(lo[:2] + 0.8, hi[:2] - 0.8)
        # keep the middle of the room free for the camera
        mid = 0.5 * (lo[:2] + hi[:2])
        if np.linalg.norm(c - mid) < 1.6:
            c = mid + (c - mid) / max(np.linalg.norm(c - mid), 1e-6) * 1.9
        bmin = np.array([c[0] - s[0] / 2, c[1] - s[1] / 2, lo[2]])
        bmax = np.array([c[0] + s[0] / 2, c[1] + s[1] / 2, lo[2] + s[2]])
        boxes.append((np.maximum(bmin, lo + 0.05), np.minimum(bmax, hi - 0.05)))
    return Scene(lo, hi, boxes)


def building_scene(rooms_x: int = 5, rooms_y: int = 4, room=(3.2, 3.0, 2.6), seed: int = 3) -> Scene:
    """~rooms_x*rooms_y*room_area m^2 floor plan: one big hall with a lattice of thick wall pillars and
    furniture boxes (walls are solid boxes with door gaps, so every surface stays closed-form)."""
    rng = np.random.default_rng(seed)
    lo = np.array([0.013, 0.017, 0.011])
    size = np.array([rooms_x * room[0], rooms_y * room[1], room[2]])
    hi = lo + size
    boxes = []
    th = 0.12
    for i in range(1, rooms_x):          # walls parallel to y with a door gap per room
        x = lo[0] + i * room[0]
        for j in range(rooms_y):
            y0, y1 = lo[1] + j * room[1], lo[1] + (j + 1) * room[1]
            d = rng.uniform(y0 + 0.5, y1 - 1.4)
            boxes.append((np.array([x - th / 2, y0, lo[2]]), np.array([x + th / 2, d, hi[2]])))
            boxes.append((np.array([x - th / 2, d + 0.9, lo[2]]), np.array([x + th / 2, y1, hi[2]])))
    for j in range(1, rooms_y):          # walls parallel to x
        y = lo[1] + j * room[1]
        for i in range(rooms_x):
            x0, x1 = lo[0] + i * room[0], lo[0] + (i + 1) * room[0]
            d = rng.uniform(x0 + 0.5, x1 - 1.4)
            boxes.append((np.array([x0, y - th / 2, lo[2]]), np.array([d, y + th / 2, hi[2]])))
            boxes.append((np.array([d + 0.9, y - th / 2, lo[2]]), np.array([x1, y + th / 2, hi[2]])))
    for i in range(rooms_x):             # one piece of furniture per room, in a corner
        for j in range(rooms_y):
            s = rng.uniform([0.4, 0.4, 0.4], [0.9, 0.9, 1.1])
            c = np.array([lo[0] + i * room[0] + 0.25 + s[0] / 2, lo[1] + j * room[1] + 0.25 + s[1] / 2])
            boxes.append((np.array([c[0] - s[0] / 2, c[1] - s[1] / 2, lo[2]]),
                          np.array([c[0] + s[0] / 2, c[1] + s[1] / 2, lo[2] + s[2]])))
    return Scene(lo, hi, boxes)


def look_at(eye, forward, up=(0.0, 0.0, 1.0)) -> SE3:
    """cam_T_world of a camera at `eye` looking along `forward` (x right, y down, z forward)."""
    f = np.asarray(forward, dtype=np.float64)
    f = f / np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, dtype=np.float64))
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    R_cw = np.stack([r, d, f])
    m = np.eye(4)
    m[:3, :3] = R_cw
    m[:3, 3] = -R_cw @ np.asarray(eye, dtype=np.float64)
    return SE3.from_matrix(m)


def circle_trajectory(scene: Scene, n_frames: int = 100, radius: float = 1.0, height: float = 1.5,
                      pitch: float = -0.15):
    """Config 1: camera on a circle around the room centre, looking outward, one turn."""
    c = 0.5 * (scene.room_min + scene.room_max)
    poses = []
    for i in range(n_frames):
        a = 2 * np.pi * i / n_frames
        eye = np.array([c[0] + radius * np.cos(a), c[1] + radius * np.sin(a), scene.room_min[2] + height])
        fwd = np.array([np.cos(a), np.sin(a), pitch])
        poses.append(look_at(eye, fwd))
    return poses


def handheld_trajectory(scene: Scene, n_frames: int = 300, seed: int = 1, speed: float = 0.012,
                        margin: float = 0.9):
    """Config 2: smoothed random walk (position + heading), TUM-fr1-like hand-held motion."""
    rng = np.random.default_rng(seed)
    c = 0.5 * (scene.room_min + scene.room_max)
    pos = np.array([c[0], c[1], scene.room_min[2] + 1.4])
    yaw, pitch = rng.uniform(0, 2 * np.pi), -0.2
    vel = np.zeros(3)
    wyaw = 0.0
    poses = []
    for _ in range(n_frames):
        vel = 0.9 * vel + 0.1 * rng.normal(0, speed, 3) * np.array([1, 1, 0.3])
        wyaw = 0.92 * wyaw + 0.08 * rng.normal(0, 0.04)
        pitch = np.clip(0.97 * pitch + 0.03 * rng.normal(-0.2, 0.25), -0.7, 0.3)
        pos = pos + vel
        lo, hi = scene.room_min + margin, scene.room_max - margin
        for k in range(3):
            if pos[k] < lo[k] or pos[k] > hi[k]:
                vel[k] = -vel[k]
                pos[k] = np.clip(pos[k], lo[k], hi[k])
        yaw += wyaw + 0.01
        fwd = np.array([np.cos(yaw) * np.cos(pitch), np.sin(yaw) * np.cos(pitch), np.sin(pitch)])
        poses.append(look_at(pos.copy(), fwd))
    return poses


def walk_trajectory(scene: Scene, n_frames: int = 200, seed: int = 3, height: float = 1.5):
    """Config 5: walk a lawnmower path through the building, panning the camera."""
    lo, hi = scene.room_min, scene.room_max
    rng = np.random.default_rng(seed)
    ys = np.linspace(lo[1] + 1.5, hi[1] - 1.5, 4)
    way = []
    for k, y in enumerate(ys):
        xs = (lo[0] + 1.5, hi[0] - 1.5) if k % 2 == 0 else (hi[0] - 1.5, lo[0] + 1.5)
        way += [np.array([xs[0], y]), np.array([xs[1], y])]
    way = np.array(way)
    seg = np.linalg.norm(np.diff(way, axis=0), axis=1)
    s = np.concatenate([[0], np.cumsum(seg)])
    poses = []
    for i in range(n_frames):
        u = s[-1] * i / max(n_frames - 1, 1)
        j = min(np.searchsorted(s, u, side="right") - 1, len(seg) - 1)
        p = way[j] + (way[j + 1] - way[j]) * ((u - s[j]) / seg[j])
        yaw = 2 * np.pi * i / 37.0 + rng.normal(0, 0.01)
        fwd = np.array([np.cos(yaw), np.sin(yaw), -0.15])
        poses.append(look_at(np.array([p[0], p[1], lo[2] + height]), fwd))
    return poses


def render(scene: Scene, K: CameraIntrinsics, H: int, W: int, cam_T_world: SE3, *, depth_factor: float | None = None,
           max_range: float = 60.0):
    """One frame: (rgb HxWx3 u8, depth HxW f32 [m], ht HxW f32, lt HxW f32).

    depth_factor: quantise through uint16 / factor like the sensors (TUM 5000, ScanNet 1000, L515 4000).
    """
    R_cw = cam_T_world.rotation_matrix()
    t_cw = cam_T_world.t.astype(np.float64)
    R_wc = R_cw.T
    eye = -R_wc @ t_cw
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    dc = np.stack([(xs - K.cx) / K.fx, (ys - K.cy) / K.fy, np.ones_like(xs)], axis=-1)   # z component 1 => t = z-depth
    dw = dc @ R_wc.T
    dw = np.where(np.abs(dw) < 1e-12, 1e-12, dw)      # axis-parallel rays: keep 1/dw finite
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dw
        # room: the camera is inside, take the nearest exit plane
        t_lo = (scene.room_min - eye) * inv
        t_hi = (scene.room_max - eye) * inv
        t_exit = np.where(dw > 0, t_hi, t_lo)
        axis_room = np.argmin(t_exit, axis=-1)
        t_room = np.min(t_exit, axis=-1)
        best = t_room.copy()
        kind = np.zeros((H, W), dtype=np.int32)          # 0 room, 1 solid box
        axis = axis_room.copy()
        side_hi = np.take_along_axis(dw > 0, axis_room[..., None], axis=-1)[..., 0]
        for bmin, bmax in scene.boxes:
            t0 = (bmin - eye) * inv
            t1 = (bmax - eye) * inv
            tn = np.minimum(t0, t1)
            tf = np.maximum(t0, t1)
            t_in = np.max(tn, axis=-1)
            t_out = np.min(tf, axis=-1)
            hit = (t_in < t_out) & (t_in > 1e-6) & (t_in < best)
            best = np.where(hit, t_in, best)
            kind = np.where(hit, 1, kind)
            axis = np.where(hit, np.argmax(tn, axis=-1), axis)
    depth = np.where(best < max_range, best, 0.0)
    pw = eye + dw * best[..., None]
    # 0.25 m checker on world coordinates (nudged inside the surface normal so it is stable)
    cell = np.floor((pw + 1e-4) / 0.25).astype(np.int64)
    chk = (cell.sum(axis=-1) & 1).astype(bool)
    base = np.where(chk[..., None], np.array([200, 60, 60]), np.array([60, 80, 200]))
    shade = (40 * (axis[..., None] == np.arange(3))).sum(axis=-1)
    rgb = np.clip(base + shade[..., None] - 20, 0, 255).astype(np.uint8)
    touch = (kind == 1) | ((kind == 0) & (axis == scene.touch_wall_axis) & side_hi)
    ht = np.where(touch, 0.9, 0.1).astype(np.float32)
    lt = (1.0 - ht).astype(np.float32)
    depth = depth.astype(np.float32)
    if depth_factor:
        q = np.clip(np.round(depth.astype(np.float64) * depth_factor), 0, 65535).astype(np.uint16)
        depth = (q.astype(np.float32) * np.float32(1.0 / depth_factor)).astype(np.float32)  # convertTo(CV_32FC1, 1/factor)
    return np.ascontiguousarray(rgb), np.ascontiguousarray(depth), ht, lt


@dataclass
class Sequence:
    name: str
    K: CameraIntrinsics
    H: int
    W: int
    poses: list
    scene: Scene
    voxel_size: float
    truncation: float
    max_depth: float
    depth_factor: float | None = None

    def frame(self, i: int):
        return render(self.scene, self.K, self.H, self.W, self.poses[i], depth_factor=self.depth_factor)

    def __len__(self):
        return len(self.poses)


def config1(n_frames: int = 100, H: int = 480, W: int = 640, truncation: float = 0.04) -> Sequence:
    """BASELINE.json configs[0]: analytic room, 640x480, voxel 0.01, trunc 0.04, max_depth 4, seed 0."""
    sc = room_scene(seed=0)
    K = TUM_FR1 if (H, W) == (480, 640) else scaled(TUM_FR1, W / 640.0, H / 480.0)
    return Sequence("analytic-room", K, H, W, circle_trajectory(sc, n_frames), sc, 0.01, truncation, 4.0)


def config2(n_frames: int = 300, H: int = 480, W: int = 640) -> Sequence:
    """configs[1]: TUM fr1-shaped, 30 Hz hand-held walk, depth through uint16/5000, trunc 6 voxels."""
    sc = room_scene(size=(5.0, 4.0, 2.8), n_boxes=4, seed=1)
    K = TUM_FR1 if (H, W) == (480, 640) else scaled(TUM_FR1, W / 640.0, H / 480.0)
    return Sequence("tum-fr1-shaped", K, H, W, handheld_trajectory(sc, n_frames, seed=1), sc, 0.01, 0.06, 4.0, 5000.0)


def config3(n_frames: int = 200, hires: bool = False) -> Sequence:
    """configs[2]: ScanNet scene0-shaped; 640x480 (what the reference runs) or 1296x968."""
    sc = room_scene(size=(7.0, 5.0, 2.8), n_boxes=5, seed=2)
    if hires:
        K, H, W = scaled(SCANNET, 1296 / 640.0, 968 / 480.0), 968, 1296
    else:
        K, H, W = SCANNET, 480, 640
    return Sequence("scannet-shaped" + ("-1296x968" if hires else ""), K, H, W,
                    handheld_trajectory(sc, n_frames, seed=2), sc, 0.01, 0.06, 4.0, 1000.0)


def config4(seed: int = 10, n_frames: int = 200, hires: bool = False) -> Sequence:
    """configs[3]: L515-shaped stream, one independent sequence per GPU (seeds 10..17)."""
    sc = room_scene(size=(6.0, 5.0, 2.8), n_boxes=4, seed=seed)
    if hires:
        K, H, W = scaled(L515, 1024 / 640.0, 768 / 480.0), 768, 1024
    else:
        K, H, W = L515, 480, 640
    return Sequence(f"l515-shaped-seed{seed}", K, H, W, handheld_trajectory(sc, n_frames, seed=seed), sc, 0.01, 0.06,
                    4.0, 4000.0)


def config5(n_frames: int = 200, H: int = 480, W: int = 640) -> Sequence:
    """configs[4]: ~200 m^2 multi-room building, 5 mm voxels, trunc 0.03, max_depth 8."""
    sc = building_scene(seed=3)
    K = TUM_FR1 if (H, W) == (480, 640) else scaled(TUM_FR1, W / 640.0, H / 480.0)
    return Sequence("building-5mm", K, H, W, walk_trajectory(sc, n_frames, seed=3), sc, 0.005, 0.03, 8.0)


def render_torch(scene: Scene, K: CameraIntrinsics, H: int, W: int, cam_T_world: SE3, device, *,
                 depth_factor: float | None = None, max_range: float = 60.0):
    """Same image formation as :func:`render`, evaluated with torch on `device` (float64).  Used by the
    large-scene benchmark where the building has ~70 boxes and numpy would take seconds per frame.
    Returns torch tensors (rgb u8 HxWx3, depth f32, ht f32, lt f32) on `device`."""
    import torch
    f64 = torch.float64
    R_cw = torch.as_tensor(cam_T_world.rotation_matrix(), dtype=f64, device=device)
    t_cw = torch.as_tensor(cam_T_world.t.astype(np.float64), dtype=f64, device=device)
    R_wc = R_cw.T
    eye = -(R_wc @ t_cw)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=f64, device=device), torch.arange(W, dtype=f64, device=device), indexing="ij")
    dc = torch.stack([(xs - K.cx) / K.fx, (ys - K.cy) / K.fy, torch.ones_like(xs)], dim=-1)
    dw = dc @ R_wc.T
    dw = torch.where(dw.abs() < 1e-12, torch.full_like(dw, 1e-12), dw)
    inv = 1.0 / dw
    rmin = torch.as_tensor(scene.room_min, dtype=f64, device=device)
    rmax = torch.as_tensor(scene.room_max, dtype=f64, device=device)
    t_exit = torch.where(dw > 0, (rmax - eye) * inv, (rmin - eye) * inv)
    best, axis = t_exit.min(dim=-1)
    side_hi = torch.gather(dw > 0, -1, axis[..., None])[..., 0]
    kind = torch.zeros((H, W), dtype=torch.int32, device=device)
    if scene.boxes:
        bmin = torch.as_tensor(np.stack([b[0] for b in scene.boxes]), dtype=f64, device=device)   # (B,3)
        bmax = torch.as_tensor(np.stack([b[1] for b in scene.boxes]), dtype=f64, device=device)
        for s in range(0, len(scene.boxes), 16):
            t0 = (bmin[s:s + 16, None, None, :] - eye) * inv
            t1 = (bmax[s:s + 16, None, None, :] - eye) * inv
            tn, tf = torch.minimum(t0, t1), torch.maximum(t0, t1)
            t_in, ax_in = tn.max(dim=-1)
            t_out = tf.min(dim=-1).values
            t_in = torch.where((t_in < t_out) & (t_in > 1e-6), t_in, torch.full_like(t_in, float("inf")))
            tb, ib = t_in.min(dim=0)
            axb = torch.gather(ax_in, 0, ib[None])[0]
            hit = tb < best
            best = torch.where(hit, tb, best)
            kind = torch.where(hit, torch.ones_like(kind), kind)
            axis = torch.where(hit, axb, axis)
    depth = torch.where(best < max_range, best, torch.zeros_like(best))
    pw = eye + dw * best[..., None]
    cell = torch.floor((pw + 1e-4) / 0.25).to(torch.int64)
    chk = (cell.sum(dim=-1) & 1).bool()
    base = torch.where(chk[..., None], torch.tensor([200, 60, 60], device=device), torch.tensor([60, 80, 200], device=device))
    shade = 40 * (axis[..., None] == torch.arange(3, device=device)).sum(dim=-1)
    rgb = (base + shade[..., None] - 20).clamp(0, 255).to(torch.uint8)
    touch = (kind == 1) | ((kind == 0) & (axis == scene.touch_wall_axis) & side_hi)
    ht = torch.where(touch, 0.9, 0.1).to(torch.float32)
    lt = (1.0 - ht).to(torch.float32)
    depth = depth.to(torch.float32)
    if depth_factor:
        q = torch.round(depth.to(f64) * depth_factor).clamp(0, 65535)
        depth = (q.to(torch.float32) * np.float32(1.0 / depth_factor)).to(torch.float32)
    return rgb.contiguous(), depth.contiguous(), ht.contiguous(), lt.contiguous()


def class_planes(ht: np.ndarray, rgb: np.ndarray, depth: np.ndarray, num_classes: int = 20, seed: int = 2) -> np.ndarray:
    """K per-class probability planes for the K-class extension: every image region (touch
    surface x checker colour x coarse depth band) gets a fixed Dirichlet(alpha = 0.3) class distribution drawn from
    `seed`; returns float32 [K, H, W] with rows summing to 1 and every entry >= 1e-4."""
    rng = np.random.default_rng(seed)
    table = rng.dirichlet(np.full(num_classes, 0.3), size=64).astype(np.float64)
    table = np.maximum(table, 1e-4)
    table /= table.sum(axis=1, keepdims=True)
    region = ((ht > 0.5).astype(np.int64) | ((rgb[..., 0] > 128).astype(np.int64) << 1) |
              (np.clip((depth * 2.0).astype(np.int64), 0, 15) << 2))
    return np.ascontiguousarray(table[region].transpose(2, 0, 1).astype(np.float32))
