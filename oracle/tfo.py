"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/libtfo.so, and the
reference-backed hybrid oracle/_ref/libtfo_ref.so when it has been built).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package topfusion_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "libtfo.so")
REF_LIB = os.path.join(_HERE, "_ref", "libtfo_ref.so")


class Params(C.Structure):
    _fields_ = [
        ("cols", C.c_int32), ("rows", C.c_int32),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("bilateral_sigma_depth", C.c_float), ("bilateral_sigma_spatial", C.c_float),
        ("bilateral_kernel_size", C.c_int32),
        ("icp_truncate_depth_dist", C.c_float), ("icp_dist_thres", C.c_float), ("icp_angle_thres", C.c_float),
        ("icp_iters", C.c_int32 * 4),
        ("mu", C.c_float), ("max_w", C.c_int32), ("voxel_size", C.c_float),
        ("view_frustum_min", C.c_float), ("view_frustum_max", C.c_float),
        ("stop_integrating_at_max_w", C.c_int32),
        ("num_blocks", C.c_int32), ("num_buckets", C.c_int32), ("excess_size", C.c_int32),
        ("depth_cutoff_mm", C.c_int32), ("corrected_mode", C.c_int32),
        ("shard_rank", C.c_int32), ("shard_count", C.c_int32),
    ]


HASH_DTYPE = np.dtype([("pos", np.int16, 3), ("pad", np.int16), ("offset", np.int32), ("ptr", np.int32)])
VOXEL_DTYPE = np.dtype([("sdf", np.int16), ("w", np.uint8), ("pad", np.uint8)])


def build(force: bool = False) -> None:
    """compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(PORT_LIB) or os.path.isdir("/root/reference/tfusion/include"):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, capture_output=True)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Lib:
    def __init__(self, which: str = "port"):
        path = PORT_LIB if which == "port" else REF_LIB
        if not os.path.exists(path):
            if which == "port":
                build()
            else:
                raise FileNotFoundError(path)
        self.which = which
        self.lib = C.CDLL(path)
        L = self.lib
        L.tfo_impl_name.restype = C.c_char_p
        L.tfo_create.restype = C.c_void_p
        L.tfo_det6.restype = C.c_double
        L.tfo_voxel_updates.restype = C.c_longlong
        for name in ("tfo_destroy", "tfo_reset", "tfo_allocate", "tfo_integrate", "tfo_expected_depths", "tfo_icp_maps",
                     "tfo_raycast", "tfo_render_image", "tfo_process_frame", "tfo_num_poses", "tfo_get_pose", "tfo_get_counters",
                     "tfo_voxel_updates", "tfo_total_entries", "tfo_export_table", "tfo_export_vis_type",
                     "tfo_export_visible_ids", "tfo_export_block", "tfo_export_minmax", "tfo_export_raycast",
                     "tfo_export_dists", "tfo_export_level", "tfo_import_level", "tfo_estimate_transform", "tfo_preprocess"):
            getattr(L, name).argtypes = None

    def impl_name(self) -> str:
        return self.lib.tfo_impl_name().decode()

    def default_params(self) -> Params:
        p = Params()
        self.lib.tfo_default_params(C.byref(p))
        return p

    # ---- stateless stages -------------------------------------------------------------
    def compute_dists(self, depth, cutoff=2047):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        out = np.empty(d.shape, np.float32)
        self.lib.tfo_compute_dists(_p(d), _p(out), C.c_int(d.shape[1]), C.c_int(d.shape[0]), C.c_int(cutoff))
        return out

    def bilateral(self, depth, ksz=7, sigma_spatial=4.5, sigma_depth=0.04):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        out = np.empty_like(d)
        self.lib.tfo_bilateral(_p(d), _p(out), C.c_int(d.shape[1]), C.c_int(d.shape[0]), C.c_int(ksz),
                               C.c_float(sigma_spatial), C.c_float(sigma_depth))
        return out

    def truncate_depth(self, depth, max_dist=2.0):
        d = np.array(depth, dtype=np.uint16, copy=True)
        self.lib.tfo_truncate_depth(_p(d), C.c_int(d.shape[1]), C.c_int(d.shape[0]), C.c_float(max_dist))
        return d

    def depth_pyr(self, depth, sigma_depth=0.04):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        out = np.empty((d.shape[0] // 2, d.shape[1] // 2), np.uint16)
        self.lib.tfo_depth_pyr(_p(d), _p(out), C.c_int(d.shape[1]), C.c_int(d.shape[0]), C.c_float(sigma_depth))
        return out

    def points_normals(self, depth, intr):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        pts = np.empty(d.shape + (4,), np.float32)
        nrm = np.empty(d.shape + (4,), np.float32)
        self.lib.tfo_points_normals(_p(d), _p(pts), _p(nrm), C.c_int(d.shape[1]), C.c_int(d.shape[0]),
                                    *[C.c_float(v) for v in intr])
        return pts, nrm

    def resize_points_normals(self, pts, nrm):
        pts = _f32(pts); nrm = _f32(nrm)
        h, w = pts.shape[:2]
        po = np.empty((h // 2, w // 2, 4), np.float32)
        no = np.empty((h // 2, w // 2, 4), np.float32)
        self.lib.tfo_resize_points_normals(_p(pts), _p(nrm), _p(po), _p(no), C.c_int(w), C.c_int(h))
        return po, no

    def icp_reduce(self, intr, aff, vcurr, ncurr, vprev, nprev, dist_thres=0.1, angle_thres=30 * 0.017453293):
        vcurr, ncurr, vprev, nprev = map(_f32, (vcurr, ncurr, vprev, nprev))
        h, w = vcurr.shape[:2]
        out = np.empty(27, np.float32)
        n = C.c_int(0)
        a = _f32(aff).reshape(16)
        self.lib.tfo_icp_reduce(C.c_int(w), C.c_int(h), *[C.c_float(v) for v in intr], C.c_float(dist_thres),
                                C.c_float(angle_thres), _p(a), _p(vcurr), _p(ncurr), _p(vprev), _p(nprev), _p(out),
                                C.byref(n))
        return out, n.value

    def icp_solve_update(self, v27, aff):
        a = _f32(aff).reshape(16).copy()
        v = _f32(v27)
        ok = self.lib.tfo_icp_solve_update(_p(v), _p(a))
        return bool(ok), a.reshape(4, 4)

    def solve6(self, A, b):
        A = _f32(A).reshape(36); b = _f32(b)
        x = np.empty(6, np.float32)
        self.lib.tfo_solve6(_p(A), _p(b), _p(x))
        return x

    def det6(self, A):
        A = _f32(A).reshape(36)
        return float(self.lib.tfo_det6(_p(A)))

    def rodrigues(self, rvec, t):
        r = _f32(rvec); t = _f32(t)
        out = np.empty(16, np.float32)
        self.lib.tfo_rodrigues(_p(r), _p(t), _p(out))
        return out.reshape(4, 4)

    def pose_inv(self, m):
        a = _f32(m).reshape(16); out = np.empty(16, np.float32)
        self.lib.tfo_pose_inv(_p(a), _p(out))
        return out.reshape(4, 4)

    def pose_mul(self, a, b):
        a = _f32(a).reshape(16); b = _f32(b).reshape(16); out = np.empty(16, np.float32)
        self.lib.tfo_pose_mul(_p(a), _p(b), _p(out))
        return out.reshape(4, 4)

    def mat4_inv_colmajor(self, m):
        a = _f32(m).reshape(16); out = np.zeros(16, np.float32)
        ok = self.lib.tfo_mat4_inv_colmajor(_p(a), _p(out))
        return bool(ok), out


class Oracle:
    """one reconstruction context — mirrors tfusion::TopFu plus stage-level entry points."""

    def __init__(self, params: Params | None = None, which: str = "port", lib: Lib | None = None, **overrides):
        self.L = lib or Lib(which)
        self.params = params or self.L.default_params()
        for k, v in overrides.items():
            if k == "icp_iters":
                for i, it in enumerate(v):
                    self.params.icp_iters[i] = it
            else:
                setattr(self.params, k, v)
        self.h = C.c_void_p(self.L.lib.tfo_create(C.byref(self.params)))
        self.cols, self.rows = self.params.cols, self.params.rows

    def close(self):
        if self.h:
            self.L.lib.tfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.L.lib.tfo_reset(self.h)

    def allocate(self, pose_w2c, dists):
        self.L.lib.tfo_allocate(self.h, _p(_f32(pose_w2c).reshape(16)), _p(_f32(dists)))

    def integrate(self, pose_w2c, dists):
        self.L.lib.tfo_integrate(self.h, _p(_f32(pose_w2c).reshape(16)), _p(_f32(dists)))

    def expected_depths(self, pose_w2c):
        self.L.lib.tfo_expected_depths(self.h, _p(_f32(pose_w2c).reshape(16)))

    def icp_maps(self, pose_c2w):
        pts = np.empty((self.rows, self.cols, 4), np.float32)
        nrm = np.empty((self.rows, self.cols, 4), np.float32)
        self.L.lib.tfo_icp_maps(self.h, _p(_f32(pose_c2w).reshape(16)), _p(pts), _p(nrm))
        return pts, nrm

    def raycast(self, pose_c2w, update_visible=True):
        self.L.lib.tfo_raycast(self.h, _p(_f32(pose_c2w).reshape(16)), C.c_int(1 if update_visible else 0))
        return self.raycast_result()

    def render_image(self, pose_c2w=None):
        pose = self.pose() if pose_c2w is None else pose_c2w
        out = np.empty((self.rows, self.cols, 4), np.uint8)
        self.L.lib.tfo_render_image(self.h, _p(_f32(pose).reshape(16)), _p(out))
        return out

    def render_point_cloud(self, pose_c2w=None, skip_points=False):
        pose = self.pose() if pose_c2w is None else pose_c2w
        out = np.empty((self.rows * self.cols, 4), np.float32)
        self.L.lib.tfo_render_point_cloud.restype = C.c_int
        n = self.L.lib.tfo_render_point_cloud(self.h, _p(_f32(pose).reshape(16)), C.c_int(1 if skip_points else 0), _p(out))
        return out[:n].copy()

    def process_frame(self, depth) -> bool:
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        assert d.shape == (self.rows, self.cols)
        return bool(self.L.lib.tfo_process_frame(self.h, _p(d)))

    def preprocess(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        self.L.lib.tfo_preprocess(self.h, _p(d))

    def estimate_transform(self):
        a = np.empty(16, np.float32)
        ok = self.L.lib.tfo_estimate_transform(self.h, _p(a))
        return bool(ok), a.reshape(4, 4)

    def num_poses(self) -> int:
        return int(self.L.lib.tfo_num_poses(self.h))

    def pose(self, idx: int = -1):
        out = np.empty(16, np.float32)
        self.L.lib.tfo_get_pose(self.h, C.c_int(idx), _p(out))
        return out.reshape(4, 4)

    def counters(self) -> dict:
        c = np.zeros(8, np.int64)
        self.L.lib.tfo_get_counters(self.h, _p(c))
        keys = ["n_visible", "last_free_block", "last_free_excess", "n_tiles", "frame_counter", "resets",
                "icp_corresp_last", "n_allocated"]
        return dict(zip(keys, (int(v) for v in c)))

    def voxel_updates(self) -> int:
        return int(self.L.lib.tfo_voxel_updates(self.h))

    def table(self):
        n = int(self.L.lib.tfo_total_entries(self.h))
        t = np.empty(n, HASH_DTYPE)
        self.L.lib.tfo_export_table(self.h, _p(t))
        return t

    def vis_type(self):
        n = int(self.L.lib.tfo_total_entries(self.h))
        t = np.empty(n, np.uint8)
        self.L.lib.tfo_export_vis_type(self.h, _p(t))
        return t

    def visible_ids(self):
        n = self.counters()["n_visible"]
        ids = np.empty(max(n, 1), np.int32)
        self.L.lib.tfo_export_visible_ids(self.h, _p(ids))
        return ids[:n]

    def block(self, ptr: int):
        b = np.empty(512, VOXEL_DTYPE)
        self.L.lib.tfo_export_block(self.h, C.c_int(ptr), _p(b))
        return b

    def blocks_by_pos(self) -> dict:
        """{(bx,by,bz): voxels[512]} for every allocated block."""
        t = self.table()
        out = {}
        for e in t[t["ptr"] >= 0]:
            out[tuple(int(v) for v in e["pos"])] = self.block(int(e["ptr"]))
        return out

    def minmax(self):
        m = np.empty((self.rows, self.cols, 2), np.float32)
        self.L.lib.tfo_export_minmax(self.h, _p(m))
        return m

    def raycast_result(self):
        m = np.empty((self.rows, self.cols, 4), np.float32)
        self.L.lib.tfo_export_raycast(self.h, _p(m))
        return m

    def dists(self):
        m = np.empty((self.rows, self.cols), np.float32)
        self.L.lib.tfo_export_dists(self.h, _p(m))
        return m

    def level(self, which: int, level: int):
        w, h = self.cols >> level, self.rows >> level
        if which == 0:
            out = np.empty((h, w), np.uint16)
        else:
            out = np.empty((h, w, 4), np.float32)
        self.L.lib.tfo_export_level(self.h, C.c_int(which), C.c_int(level), _p(out))
        return out

    def set_level(self, which: int, level: int, arr):
        a = np.ascontiguousarray(arr, dtype=np.uint16 if which == 0 else np.float32)
        self.L.lib.tfo_import_level(self.h, C.c_int(which), C.c_int(level), _p(a))


def allocated_set(table) -> set:
    t = table[table["ptr"] >= -1]
    return {tuple(int(v) for v in e) for e in t["pos"]}


def visible_set(table, ids) -> set:
    """visible list as a set of allocated block coordinates."""
    e = table[ids]
    e = e[e["ptr"] >= -1]
    return {tuple(int(v) for v in p) for p in e["pos"]}
