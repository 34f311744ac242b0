"""Synthetic depth sequences for the parity tests and the bench.

The reference ships no sample data (its demo reads ``%04d.pgm`` frames from a
hard-coded path, /root/reference/apps/demo.cpp:91-100), so every input here is
an analytic pinhole rendering of simple solids: z-depth in millimetres rounded
to u16, 0 = no return.  Deterministic for a given (sequence, frame index).

Sequences
  S0 "hover"  S1 geometry, camera jitters around the origin (<= 5 mm, <= 0.2 deg)
  S1 "orbit"  sphere + box + floor, camera orbits (0,0,1.2) by 0.5 deg/frame
  S2 "room"   1280x720 box room seen from inside, walk 1 cm + 0.3 deg/frame
  S3 "large"  2x1x2 m room shell + 20 seeded boxes (2 mm voxels in the configs)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# reference defaults, /root/reference/tfusion/src/topfu.cpp:24
DEFAULT_INTR = (504.261, 503.905, 352.457, 272.202)


@dataclass
class Scene:
    spheres: list = field(default_factory=list)   # (cx, cy, cz, r)
    boxes: list = field(default_factory=list)     # (xmin, ymin, zmin, xmax, ymax, zmax) hit from outside
    planes: list = field(default_factory=list)    # (nx, ny, nz, d): n.p = d, hit when n.dir > 0
    rooms: list = field(default_factory=list)     # boxes seen from the inside


def scene_s1() -> Scene:
    s = Scene()
    s.spheres.append((0.0, 0.0, 1.2, 0.35))
    cx, cy, cz, sx, sy, sz = 0.45, 0.1, 1.3, 0.5, 0.3, 0.4
    s.boxes.append((cx - sx / 2, cy - sy / 2, cz - sz / 2, cx + sx / 2, cy + sy / 2, cz + sz / 2))
    s.planes.append((0.0, 1.0, 0.0, 0.45))  # floor y = 0.45 (y points down)
    return s


def scene_s2() -> Scene:
    s = Scene()
    s.rooms.append((-2.0, -1.25, -0.5, 2.0, 1.25, 3.5))
    s.boxes.append((-1.2, 0.45, 1.6, -0.4, 1.25, 2.4))
    s.boxes.append((0.5, 0.25, 2.0, 1.3, 1.25, 2.6))
    return s


def scene_s3(seed: int = 7) -> Scene:
    s = Scene()
    s.rooms.append((-1.0, -0.5, 0.2, 1.0, 0.5, 2.2))
    rng = np.random.RandomState(seed)
    for _ in range(20):
        c = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.3, 0.4), rng.uniform(0.9, 2.0)])
        h = rng.uniform(0.04, 0.14, size=3)
        s.boxes.append(tuple(np.concatenate([c - h, c + h])))
    return s


def _rot_y(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_x(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _pose(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    m = np.eye(4, dtype=np.float64)
    m[:3, :3] = R
    m[:3, 3] = t
    return m


def pose_orbit(i: int, deg_per_frame: float = 0.5, pivot=(0.0, 0.0, 1.2)) -> np.ndarray:
    """camera->world pose of orbit frame i (frame 0 = identity)."""
    a = math.radians(deg_per_frame) * i
    R = _rot_y(a)
    p = np.asarray(pivot, dtype=np.float64)
    t = p + R @ (-p)
    return _pose(R, t)


def pose_hover(i: int, seed: int = 42) -> np.ndarray:
    """smooth pseudo-random jitter: <= 5 mm translation, <= 0.2 deg rotation; frame 0 = identity."""
    rng = np.random.RandomState(seed)
    amp_t = rng.uniform(0.0008, 0.0016, size=(3, 2))
    amp_r = np.radians(rng.uniform(0.03, 0.06, size=(2, 2)))
    frq = rng.uniform(0.04, 0.11, size=(5, 2))
    s = lambda k, j: math.sin(2 * math.pi * frq[k, j] * i)
    t = np.array([amp_t[k, 0] * s(k, 0) + amp_t[k, 1] * s(k, 1) for k in range(3)])
    rx = amp_r[0, 0] * s(3, 0) + amp_r[0, 1] * s(3, 1)
    ry = amp_r[1, 0] * s(4, 0) + amp_r[1, 1] * s(4, 1)
    return _pose(_rot_y(ry) @ _rot_x(rx), t)


def pose_walk(i: int) -> np.ndarray:
    """S2/S3 walk: 1 cm forward-ish drift plus 0.3 deg/frame yaw sweep (bounded)."""
    yaw = math.radians(25.0) * math.sin(2 * math.pi * i * 0.3 / 100.0)
    t = np.array([0.25 * math.sin(2 * math.pi * i / 157.0), 0.0, 0.4 * (1 - math.cos(2 * math.pi * i / 251.0)) * 0.5])
    return _pose(_rot_y(yaw), t)


def intrinsics_for(cols: int, rows: int):
    sx = cols / 640.0
    sy = rows / 480.0
    fx, fy, cx, cy = DEFAULT_INTR
    if (cols, rows) == (640, 480):
        return DEFAULT_INTR
    s = sx if abs(sx - sy) < 1e-9 else min(sx, sy)
    return (fx * s, fy * s, cx * sx, cy * sy)


def render_depth(scene: Scene, pose_c2w: np.ndarray, cols: int = 640, rows: int = 480,
                 intr=None, noise_mm: float = 0.0, seed: int = 1234,
                 max_mm: int = 10000) -> np.ndarray:
    """z-depth in mm (u16), 0 = miss.  Ray parameter t equals camera z because dir_cam.z = 1.

    Returns beyond max_mm (10 m, a depth sensor's range) are dropped: the reference's bilateral kernel squares
    (value - depth) in int32 (imgproc.cu:36), which overflows for differences above 46 340 mm."""
    fx, fy, cx, cy = intr if intr is not None else intrinsics_for(cols, rows)
    u = (np.arange(cols, dtype=np.float64) - cx) / fx
    v = (np.arange(rows, dtype=np.float64) - cy) / fy
    dc = np.stack(np.broadcast_arrays(u[None, :], v[:, None], np.ones((rows, cols))), axis=-1)
    R = pose_c2w[:3, :3]
    o = pose_c2w[:3, 3]
    d = dc @ R.T
    best = np.full((rows, cols), np.inf)
    eps = 1e-6
    for (sx_, sy_, sz_, r) in scene.spheres:
        oc = o - np.array([sx_, sy_, sz_])
        a = np.sum(d * d, axis=-1)
        b = 2.0 * (d @ oc)
        c = float(oc @ oc) - r * r
        disc = b * b - 4 * a * c
        ok = disc >= 0
        sq = np.sqrt(np.where(ok, disc, 0.0))
        t0 = (-b - sq) / (2 * a)
        t1 = (-b + sq) / (2 * a)
        t = np.where(t0 > eps, t0, t1)
        t = np.where(ok & (t > eps), t, np.inf)
        best = np.minimum(best, t)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        for bx in list(scene.boxes) + list(scene.rooms):
            lo = (np.array(bx[:3]) - o) * inv
            hi = (np.array(bx[3:]) - o) * inv
            tn = np.minimum(lo, hi).max(axis=-1)
            tf = np.maximum(lo, hi).min(axis=-1)
            hit = tf >= np.maximum(tn, 0.0)
            if bx in scene.rooms:
                t = np.where(hit & (tf > eps), tf, np.inf)      # far wall from inside
            else:
                t = np.where(hit & (tn > eps), tn, np.inf)
            best = np.minimum(best, t)
        for (nx, ny, nz, dd) in scene.planes:
            n = np.array([nx, ny, nz])
            den = d @ n
            t = (dd - float(o @ n)) / den
            t = np.where((den > 1e-9) & (t > eps), t, np.inf)
            best = np.minimum(best, t)
    mm = best * 1000.0
    if noise_mm > 0:
        rng = np.random.RandomState(seed)
        mm = mm + rng.normal(0.0, noise_mm, size=mm.shape)
    mm = np.where(np.isfinite(mm), np.rint(mm), 0.0)
    mm = np.where((mm > 0) & (mm <= max_mm), mm, 0.0)
    return mm.astype(np.uint16)


_SEQ = {
    "S0": (scene_s1, pose_hover, 640, 480),
    "S1": (scene_s1, pose_orbit, 640, 480),
    "S2": (scene_s2, pose_walk, 1280, 720),
    "S3": (scene_s3, pose_walk, 640, 480),
}


def sequence(name: str, n_frames: int, cols: int | None = None, rows: int | None = None):
    """returns (depth[n,rows,cols] u16, poses_c2w[n,4,4] f64, intr)"""
    mk_scene, mk_pose, c0, r0 = _SEQ[name]
    cols = cols or c0
    rows = rows or r0
    sc = mk_scene()
    intr = intrinsics_for(cols, rows)
    depth = np.empty((n_frames, rows, cols), dtype=np.uint16)
    poses = np.empty((n_frames, 4, 4), dtype=np.float64)
    for i in range(n_frames):
        poses[i] = mk_pose(i)
        depth[i] = render_depth(sc, poses[i], cols, rows, intr)
    return depth, poses, intr


def write_pgm(path: str, depth: np.ndarray) -> None:
    """16-bit binary PGM, the format demo.cpp reads with cv::imread(..., CV_16U)."""
    h, w = depth.shape
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n65535\n" % (w, h))
        f.write(depth.astype(">u2").tobytes())
