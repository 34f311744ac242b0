"""REFERENCE-ARM INFRASTRUCTURE (not product code).  ctypes view of baseline/_ref/libref_gpu*.so — the reference
library's own GPU path, patched to compile (baseline/ref_gpu/patch_ref.py, Makefile).  Used by
  * tests/golden/make_refgpu_golden.py  (fixtures from the reference's device kernels, run on the B200),
  * tests/test_gpu_vs_reference_gpu.py  (live comparison when the library travelled to the GPU box),
  * bench.py's `reference_gpu` object   (GPU-vs-GPU anchor).
Nothing under topfusion_b200/, src/ or include/ imports it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
DEFAULT_INTR = (504.261, 503.905, 352.457, 272.202)
HASH_DTYPE = np.dtype([("pos", np.int16, 3), ("pad", np.int16), ("offset", np.int32), ("ptr", np.int32)])
VOXEL_DTYPE = np.dtype([("sdf", np.int16), ("w", np.uint8), ("pad", np.uint8)])


def lib_path(variant: str = "nodebug") -> str:
    return os.path.join(OUT, "libref_gpu_nodebug.so" if variant == "nodebug" else "libref_gpu.so")


def available(variant: str = "nodebug") -> bool:
    return os.path.exists(lib_path(variant))


def build() -> bool:
    """(Re)build where the reference tree exists; elsewhere keep whatever travelled.  Returns availability."""
    if os.path.isdir("/root/reference/tfusion/src"):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)
    return available("nodebug") and available("asis")


_libs = {}


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def lib(variant: str = "nodebug") -> C.CDLL:
    if variant in _libs:
        return _libs[variant]
    L = C.CDLL(lib_path(variant))
    L.refgpu_describe.restype = C.c_char_p
    L.refgpu_create.restype = C.c_void_p
    L.refgpu_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_int]
    for name in ("refgpu_destroy", "refgpu_upload", "refgpu_get_dists", "refgpu_get_raycast", "refgpu_get_range_image", "refgpu_get_hash",
                 "refgpu_get_voxels", "refgpu_get_counters", "refgpu_render"):
        getattr(L, name).argtypes = [C.c_void_p] + ([C.c_void_p] if name != "refgpu_destroy" else [])
        getattr(L, name).restype = None
    L.refgpu_step.argtypes = [C.c_void_p]; L.refgpu_step.restype = C.c_int
    L.refgpu_frame.argtypes = [C.c_void_p, C.c_void_p]; L.refgpu_frame.restype = C.c_int
    L.refgpu_num_poses.argtypes = [C.c_void_p]; L.refgpu_num_poses.restype = C.c_int
    L.refgpu_frame_counter.argtypes = [C.c_void_p]; L.refgpu_frame_counter.restype = C.c_int
    L.refgpu_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]; L.refgpu_pose.restype = None
    L.refgpu_get_maps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]; L.refgpu_get_maps.restype = None
    L.refgpu_get_depth_pyr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]; L.refgpu_get_depth_pyr.restype = None
    L.refgpu_get_visible.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; L.refgpu_get_visible.restype = C.c_int
    L.refgpu_stage_imgproc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.refgpu_stage_imgproc.restype = None
    L.refgpu_stage_resize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]; L.refgpu_stage_resize.restype = None
    L.refgpu_stage_icp_sums.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.refgpu_stage_icp_sums.restype = None
    L.refgpu_stage_estimate.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    L.refgpu_stage_estimate.restype = C.c_int
    L.refgpu_cv_determinant6.argtypes = [C.c_void_p]; L.refgpu_cv_determinant6.restype = C.c_double
    for name in ("refgpu_cv_solve6", "refgpu_cv_affine", "refgpu_cv_affine_mul"):
        getattr(L, name).argtypes = [C.c_void_p] * 3; getattr(L, name).restype = None
    L.refgpu_cv_affine_inv.argtypes = [C.c_void_p] * 2; L.refgpu_cv_affine_inv.restype = None
    L.refgpu_sync.restype = None
    L.refgpu_scene_integrate.argtypes = [C.c_void_p] * 3; L.refgpu_scene_integrate.restype = None
    L.refgpu_scene_raycast.argtypes = [C.c_void_p] * 2; L.refgpu_scene_raycast.restype = None
    L.refgpu_render_at.argtypes = [C.c_void_p] * 3; L.refgpu_render_at.restype = None
    _libs[variant] = L
    return L


ANGLE_30 = 30.0 * 0.017453293


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def stage_imgproc(depth, intr=DEFAULT_INTR, ksz=7, sigma_s=4.5, sigma_d=0.04, trunc=2.0, levels=3, variant="nodebug"):
    """the reference's imgproc.cu kernels in topfu.cpp:166-197 order -> dict of arrays"""
    L = lib(variant)
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    rows, cols = depth.shape
    out = {"dists": np.empty((rows, cols), np.float32), "bilateral": np.empty((rows, cols), np.uint16), "depth": [], "points": [], "normals": []}
    for l in range(levels):
        r, c = rows >> l, cols >> l
        out["depth"].append(np.empty((r, c), np.uint16))
        out["points"].append(np.empty((r, c, 4), np.float32))
        out["normals"].append(np.empty((r, c, 4), np.float32))
    PA = C.c_void_p * levels
    i4 = _f32(intr)
    L.refgpu_stage_imgproc(_p(depth), rows, cols, _p(i4), ksz, sigma_s, sigma_d, trunc, levels, _p(out["dists"]), _p(out["bilateral"]),
                           PA(*[a.ctypes.data for a in out["depth"]]), PA(*[a.ctypes.data for a in out["points"]]),
                           PA(*[a.ctypes.data for a in out["normals"]]))
    return out


def stage_resize(points, normals, variant="nodebug"):
    L = lib(variant)
    points, normals = _f32(points), _f32(normals)
    rows, cols = points.shape[:2]
    po = np.empty((rows // 2, cols // 2, 4), np.float32); no = np.empty_like(po)
    L.refgpu_stage_resize(_p(points), _p(normals), rows, cols, _p(po), _p(no))
    return po, no


def stage_icp_sums(intr, aff, vcurr, ncurr, vprev, nprev, level=0, dist_thres=0.1, angle_thres=ANGLE_30, variant="nodebug"):
    """icp_helper_kernel + icp_final_reduce_kernel at the fixed transform `aff` (4x4, row-major).
    `intr` are LEVEL-0 intrinsics; `level` divides them as setLevelIntr does.  Returns the 27 floats."""
    L = lib(variant)
    vcurr, ncurr, vprev, nprev = map(_f32, (vcurr, ncurr, vprev, nprev))
    rows, cols = vcurr.shape[:2]
    aff = np.asarray(aff, np.float32)
    R = _f32(aff[:3, :3]); t = _f32(aff[:3, 3]); i4 = _f32(intr)
    out = np.zeros(27, np.float32)
    L.refgpu_stage_icp_sums(_p(vcurr), _p(ncurr), _p(vprev), _p(nprev), rows, cols, level, _p(i4), _p(R), _p(t), dist_thres, angle_thres,
                            _p(out), None, None)
    return out


def stage_estimate(intr, vcurr, ncurr, vprev, nprev, iters=(10, 5, 4, 0), dist_thres=0.1, angle_thres=ANGLE_30, variant="nodebug"):
    """ProjectiveICP::estimateTransform over per-level map lists -> (ok, 4x4)"""
    L = lib(variant)
    levels = len(vcurr)
    arrs = [[_f32(a) for a in lst] for lst in (vcurr, ncurr, vprev, nprev)]
    rows, cols = arrs[0][0].shape[:2]
    PA = C.c_void_p * levels
    ptrs = [PA(*[a.ctypes.data for a in lst]) for lst in arrs]
    i4 = _f32(intr); it = np.ascontiguousarray(iters, dtype=np.int32)
    out = np.zeros((4, 4), np.float32)
    ok = L.refgpu_stage_estimate(ptrs[0], ptrs[1], ptrs[2], ptrs[3], rows, cols, levels, _p(i4), _p(it), dist_thres, angle_thres, _p(out))
    return bool(ok), out


class RefTopFu:
    """tfusion::TopFu of the reference, driven like apps/demo.cpp drives it"""

    def __init__(self, rows=480, cols=640, intr=DEFAULT_INTR, voxel=0.005, mu=0.02, max_w=100, vf_min=0.2, vf_max=3.0, iters=(10, 5, 4, 0),
                 icp_trunc=2.0, variant="nodebug", quiet=True):
        if (cols * 4) % 512:
            raise ValueError("the reference indexes pitched images densely (SURVEY F9): cols must be a multiple of 128")
        self.L = lib(variant)
        self.rows, self.cols = rows, cols
        i4 = _f32(intr); it = np.ascontiguousarray(iters, dtype=np.int32)
        self.h = self.L.refgpu_create(rows, cols, _p(i4), voxel, mu, max_w, vf_min, vf_max, _p(it), icp_trunc, 1 if quiet else 0)
        self.total = self.L.refgpu_hash_total_entries()
        self.nblocks = self.L.refgpu_num_blocks()

    def close(self):
        if self.h:
            self.L.refgpu_destroy(self.h); self.h = None

    def frame(self, depth) -> bool:
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        return bool(self.L.refgpu_frame(self.h, _p(depth)))

    def upload(self, depth):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        self.L.refgpu_upload(self.h, _p(depth))

    def step(self) -> bool:
        return bool(self.L.refgpu_step(self.h))

    def sync(self):
        self.L.refgpu_sync()

    def num_poses(self):
        return self.L.refgpu_num_poses(self.h)

    def pose(self, idx=-1):
        m = np.zeros((4, 4), np.float32)
        self.L.refgpu_pose(self.h, idx, _p(m))
        return m

    def maps(self, which, level):
        r, c = self.rows >> level, self.cols >> level
        p = np.empty((r, c, 4), np.float32); n = np.empty_like(p)
        self.L.refgpu_get_maps(self.h, which, level, _p(p), _p(n))
        return p, n

    def depth_pyr(self, level):
        a = np.empty((self.rows >> level, self.cols >> level), np.uint16)
        self.L.refgpu_get_depth_pyr(self.h, level, _p(a)); return a

    def dists(self):
        a = np.empty((self.rows, self.cols), np.float32); self.L.refgpu_get_dists(self.h, _p(a)); return a

    def raycast_result(self):
        a = np.empty((self.rows, self.cols, 4), np.float32); self.L.refgpu_get_raycast(self.h, _p(a)); return a

    def range_image(self):
        a = np.empty((self.rows, self.cols, 2), np.float32); self.L.refgpu_get_range_image(self.h, _p(a)); return a

    def table(self):
        a = np.empty(self.total, HASH_DTYPE); self.L.refgpu_get_hash(self.h, _p(a)); return a

    def voxels(self):
        a = np.empty((self.nblocks, 512), VOXEL_DTYPE); self.L.refgpu_get_voxels(self.h, _p(a)); return a

    def visible(self):
        ids = np.empty(self.nblocks, np.int32); types = np.empty(self.total, np.uint8)
        n = self.L.refgpu_get_visible(self.h, _p(ids), _p(types))
        return ids[:n].copy(), types

    def counters(self):
        a = np.zeros(3, np.int32); self.L.refgpu_get_counters(self.h, _p(a))
        return {"last_free_block": int(a[0]), "last_free_excess": int(a[1]), "n_visible": int(a[2])}

    def scene_integrate(self, depth, pose_c2w):
        """computeDists + AllocateSceneFromDepth + IntegrateIntoScene at an injected pose"""
        depth = np.ascontiguousarray(depth, dtype=np.uint16); m = _f32(pose_c2w)
        self.L.refgpu_scene_integrate(self.h, _p(depth), _p(m))

    def scene_raycast(self, pose_c2w):
        """CreateExpectedDepths + CreateICPMaps + the model-map pyramid at an injected pose"""
        m = _f32(pose_c2w)
        self.L.refgpu_scene_raycast(self.h, _p(m))

    def render_at(self, pose_c2w):
        a = np.empty((self.rows, self.cols, 4), np.uint8); m = _f32(pose_c2w)
        self.L.refgpu_render_at(self.h, _p(m), _p(a)); return a

    def render(self):
        a = np.empty((self.rows, self.cols, 4), np.uint8); self.L.refgpu_render(self.h, _p(a)); return a
