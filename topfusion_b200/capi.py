"""ctypes binding of the C ABI in include/tfusion_b200.h (topfusion_b200/libtfusion_b200.so).

This is harness-side plumbing for tests and bench.py: the product is the shared library plus the
C++ `tfusion` mirror in include/tfusion/ + src/.  There is no CPU fallback anywhere: if the
library is missing or there is no CUDA device, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFB_LIB_PATH") or os.path.join(_HERE, "libtfusion_b200.so")   # TFB_LIB_PATH: instrumented build (tools/)

HASH_DTYPE = np.dtype([("pos", np.int16, 3), ("pad", np.int16), ("offset", np.int32), ("ptr", np.int32)])
VOXEL_DTYPE = np.dtype([("sdf", np.int16), ("w", np.uint8), ("pad", np.uint8)])

STAGES = ["upload", "preprocess", "icp", "allocate", "integrate", "expected_depth", "raycast_icp_maps", "map_pyramid", "frame"]


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
        ("defer_tail", C.c_int32),
        ("ieee_arith", C.c_int32),
    ]


class ShardPtrs(C.Structure):
    """tfb_shard_ptrs: the buffers of one rank that the other ranks read / write over peer memory"""
    _fields_ = [("table", C.c_void_p), ("vba", C.c_void_p), ("raycast", C.c_void_p), ("marks", C.c_void_p),
                ("frame", C.c_void_p), ("flags", C.c_void_p)]


class TfbError(RuntimeError):
    pass


def build(verbose: bool = False) -> None:
    """compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise TfbError("nvcc build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TfbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.tfb_last_error.restype = C.c_char_p
        L.tfb_version.restype = C.c_char_p
        L.tfb_voxel_updates_last.restype = C.c_longlong
        L.tfb_voxel_updates_total.restype = C.c_longlong
        L.tfb_voxel_updates_total.argtypes = [C.c_void_p]
        L.tfb_kernel_launches.restype = C.c_longlong
        L.tfb_level_ptr.restype = C.c_void_p
        L.tfb_ktiming_name.restype = C.c_char_p
        L.tfb_stream.restype = C.c_void_p
        for name, args in {
            "tfb_create": [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)],
            "tfb_dev_alloc": [C.POINTER(C.c_void_p), C.c_size_t],
            "tfb_dev_free": [C.c_void_p],
            "tfb_host_alloc_pinned": [C.POINTER(C.c_void_p), C.c_size_t],
            "tfb_host_free_pinned": [C.c_void_p],
            "tfb_h2d": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
            "tfb_d2h": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t],
            "tfb_process_frame": [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)],
            "tfb_process_frame_device": [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)],
            "tfb_level_ptr": [C.c_void_p, C.c_int, C.c_int],
            "tfb_stream": [C.c_void_p],
            "tfb_shard_local_ptrs": [C.c_void_p, C.POINTER(ShardPtrs)],
            "tfb_shard_attach": [C.c_void_p, C.c_int, C.POINTER(ShardPtrs)],
            "tfb_ipc_export": [C.c_void_p, C.c_char_p],
            "tfb_ipc_open": [C.c_char_p, C.POINTER(C.c_void_p)],
            "tfb_ipc_close": [C.c_void_p],
            "tfb_frame_begin": [C.c_void_p, C.c_void_p],
            "tfb_frame_raycast": [C.c_void_p],
            "tfb_extract_points": [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)],
            "tfb_render_point_cloud": [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)],
            "tfb_scene_save": [C.c_void_p, C.c_char_p],
            "tfb_scene_load": [C.c_void_p, C.c_char_p],
            "tfb_stream_out": [C.c_void_p, C.c_int, C.POINTER(C.c_int)],
            "tfb_stream_in": [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
            "tfb_stream_stats": [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
            "tfb_shard_push_frame": [C.c_void_p, C.c_void_p],
            "tfb_process_frame_sharded": [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)],
            "tfb_shard_barrier": [C.c_void_p],
            "tfb_frame_end": [C.c_void_p, C.POINTER(C.c_int)],
        }.items():
            getattr(L, name).argtypes = args
        _lib = L
    return _lib


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class DevBuf:
    """a raw device allocation (cuda::DeviceMemory of the reference, src/device_memory.cpp)."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        self.nbytes = nbytes
        if lib().tfb_dev_alloc(C.byref(self.ptr), C.c_size_t(nbytes)) != 0:
            raise TfbError("device allocation failed")

    def free(self):
        if self.ptr:
            lib().tfb_dev_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """numpy view over page-locked host memory (what a capture ring would hand to process_frame)."""

    def __init__(self, shape, dtype):
        self.ptr = C.c_void_p()
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        if lib().tfb_host_alloc_pinned(C.byref(self.ptr), C.c_size_t(n)) != 0:
            raise TfbError("pinned allocation failed")
        buf = (C.c_char * n).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().tfb_host_free_pinned(self.ptr)
            self.ptr = C.c_void_p()


class Context:
    """one reconstruction context; mirrors tfusion::TopFu plus the stage-level seam."""

    def __init__(self, params: Params | None = None, stream: int | None = None, **overrides):
        L = lib()
        self.L = L
        if params is None:
            params = Params()
            L.tfb_default_params(C.byref(params))
        for k, v in overrides.items():
            if k == "icp_iters":
                for i, it in enumerate(v):
                    params.icp_iters[i] = it
            else:
                setattr(params, k, v)
        self.params = params
        self.h = C.c_void_p()
        rc = L.tfb_create(C.byref(params), C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc != 0:
            raise TfbError(f"tfb_create failed with {rc} (no CUDA device? there is no CPU fallback)")
        self.cols, self.rows = params.cols, params.rows
        self._scratch = {}

    # -- helpers -----------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise TfbError(f"tfb error {rc}: {self.L.tfb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            for b in self._scratch.values():
                b.free()
            self._scratch = {}
            self.L.tfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, arr: np.ndarray, key: str | None = None) -> DevBuf:
        a = np.ascontiguousarray(arr)
        if key is not None and key in self._scratch and self._scratch[key].nbytes >= a.nbytes:
            buf = self._scratch[key]
        else:
            buf = DevBuf(a.nbytes)
            if key is not None:
                if key in self._scratch:
                    self._scratch[key].free()
                self._scratch[key] = buf
        self._ck(self.L.tfb_h2d(self.h, buf.ptr, _np_ptr(a), C.c_size_t(a.nbytes)))
        self.sync()  # the source array may be a temporary
        return buf

    def download(self, buf_ptr, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        p = buf_ptr.ptr if isinstance(buf_ptr, DevBuf) else C.c_void_p(buf_ptr)
        self._ck(self.L.tfb_d2h(self.h, _np_ptr(out), p, C.c_size_t(out.nbytes)))
        return out

    def sync(self):
        self._ck(self.L.tfb_sync(self.h))

    def reset(self):
        self._ck(self.L.tfb_reset(self.h))

    # -- image stages (host arrays in, host arrays out; device work in between) -----------------------
    def compute_dists(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = d.shape
        src = self.upload(d, "img_in")
        dst = DevBuf(w * h * 4)
        self._ck(self.L.tfb_compute_dists(self.h, src.ptr, dst.ptr, C.c_int(w), C.c_int(h)))
        return self.download(dst, (h, w), np.float32)

    def bilateral(self, depth, ksz=7, sigma_spatial=4.5, sigma_depth=0.04):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = d.shape
        src = self.upload(d, "img_in")
        dst = DevBuf(w * h * 2)
        self._ck(self.L.tfb_bilateral_filter(self.h, src.ptr, dst.ptr, C.c_int(w), C.c_int(h), C.c_int(ksz),
                                             C.c_float(sigma_spatial), C.c_float(sigma_depth)))
        return self.download(dst, (h, w), np.uint16)

    def truncate_depth(self, depth, max_dist=2.0):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = d.shape
        src = self.upload(d, "img_in")
        self._ck(self.L.tfb_truncate_depth(self.h, src.ptr, C.c_int(w), C.c_int(h), C.c_float(max_dist)))
        return self.download(src, (h, w), np.uint16)

    def depth_pyr(self, depth, sigma_depth=0.04):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = d.shape
        src = self.upload(d, "img_in")
        dst = DevBuf((w // 2) * (h // 2) * 2)
        self._ck(self.L.tfb_depth_pyr(self.h, src.ptr, dst.ptr, C.c_int(w), C.c_int(h), C.c_float(sigma_depth)))
        return self.download(dst, (h // 2, w // 2), np.uint16)

    def points_normals(self, depth, intr):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        h, w = d.shape
        src = self.upload(d, "img_in")
        pts, nrm = DevBuf(w * h * 16), DevBuf(w * h * 16)
        self._ck(self.L.tfb_compute_point_normals(self.h, src.ptr, pts.ptr, nrm.ptr, C.c_int(w), C.c_int(h),
                                                  *[C.c_float(v) for v in intr]))
        return self.download(pts, (h, w, 4), np.float32), self.download(nrm, (h, w, 4), np.float32)

    def resize_points_normals(self, pts, nrm):
        pts, nrm = _f32(pts), _f32(nrm)
        h, w = pts.shape[:2]
        dp, dn = self.upload(pts, "map_v"), self.upload(nrm, "map_n")
        po, no = DevBuf((w // 2) * (h // 2) * 16), DevBuf((w // 2) * (h // 2) * 16)
        self._ck(self.L.tfb_resize_points_normals(self.h, dp.ptr, dn.ptr, po.ptr, no.ptr, C.c_int(w), C.c_int(h)))
        return self.download(po, (h // 2, w // 2, 4), np.float32), self.download(no, (h // 2, w // 2, 4), np.float32)

    def icp_reduce(self, intr, aff, vcurr, ncurr, vprev, nprev):
        vcurr, ncurr, vprev, nprev = map(_f32, (vcurr, ncurr, vprev, nprev))
        h, w = vcurr.shape[:2]
        bufs = [self.upload(a, k) for a, k in ((vcurr, "icp_vc"), (ncurr, "icp_nc"), (vprev, "icp_vp"), (nprev, "icp_np"))]
        a16 = _f32(aff).reshape(16)
        out = np.empty(27, np.float32)
        self._ck(self.L.tfb_icp_reduce(self.h, C.c_int(w), C.c_int(h), *[C.c_float(v) for v in intr], _np_ptr(a16),
                                       bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, _np_ptr(out)))
        return out

    # -- context-level stages ------------------------------------------------------------------------
    def preprocess(self, depth):
        d = np.ascontiguousarray(depth, dtype=np.uint16)
        src = self.upload(d, "frame_in")
        self._ck(self.L.tfb_preprocess(self.h, src.ptr))

    def estimate_transform(self):
        a = np.empty(16, np.float32)
        ok = C.c_int(0)
        self._ck(self.L.tfb_icp_estimate(self.h, _np_ptr(a), C.byref(ok)))
        return bool(ok.value), a.reshape(4, 4)

    def allocate(self, pose_w2c, dists):
        d = self.upload(_f32(dists), "dists_in")
        self._ck(self.L.tfb_allocate_scene_from_depth(self.h, _np_ptr(_f32(pose_w2c).reshape(16)), d.ptr))
        self.sync()

    def integrate(self, pose_w2c, dists):
        d = self.upload(_f32(dists), "dists_in")
        self._ck(self.L.tfb_integrate_into_scene(self.h, _np_ptr(_f32(pose_w2c).reshape(16)), d.ptr))

    def expected_depths(self, pose_w2c):
        self._ck(self.L.tfb_create_expected_depths(self.h, _np_ptr(_f32(pose_w2c).reshape(16))))
        self.sync()

    def icp_maps(self, pose_c2w):
        n = self.cols * self.rows * 16
        pts, nrm = DevBuf(n), DevBuf(n)
        self._ck(self.L.tfb_create_icp_maps(self.h, _np_ptr(_f32(pose_c2w).reshape(16)), pts.ptr, nrm.ptr))
        return (self.download(pts, (self.rows, self.cols, 4), np.float32),
                self.download(nrm, (self.rows, self.cols, 4), np.float32))

    def render_image(self, pose_c2w=None):
        """TopFu::renderImage: shaded greyscale view, uint8 [rows, cols, 4]"""
        out = DevBuf(self.rows * self.cols * 4)
        p = _np_ptr(_f32(pose_c2w).reshape(16)) if pose_c2w is not None else None
        self._ck(self.L.tfb_render_image(self.h, p, out.ptr))
        return self.download(out, (self.rows, self.cols, 4), np.uint8)

    # -- the reconstruction out and back in ------------------------------------------------------------------
    def extract_points(self) -> np.ndarray:
        """surface points of the whole scene, float32 [n, 4] (x, y, z, 1) in world metres, arbitrary order"""
        n = C.c_int(0)
        self._ck(self.L.tfb_extract_points(self.h, None, C.c_int(0), C.byref(n)))
        if n.value == 0:
            return np.zeros((0, 4), np.float32)
        buf = DevBuf(n.value * 16)
        m = C.c_int(0)
        self._ck(self.L.tfb_extract_points(self.h, buf.ptr, C.c_int(n.value), C.byref(m)))
        assert m.value == n.value
        out = self.download(buf, (n.value, 4), np.float32)
        buf.free()
        return out

    def render_point_cloud(self, pose_c2w=None, skip_points=False) -> np.ndarray:
        """the reference's renderPointCloud_device: surface points of one view, float32 [n, 4] in world metres, arbitrary order"""
        cap = self.rows * self.cols
        buf = DevBuf(cap * 16)
        n = C.c_int(0)
        p = _np_ptr(_f32(pose_c2w).reshape(16)) if pose_c2w is not None else None
        self._ck(self.L.tfb_render_point_cloud(self.h, p, C.c_int(1 if skip_points else 0), buf.ptr, C.c_int(cap), C.byref(n)))
        out = self.download(buf, (n.value, 4), np.float32) if n.value else np.zeros((0, 4), np.float32)
        buf.free()
        return out

    def save_scene(self, path: str):
        self._ck(self.L.tfb_scene_save(self.h, path.encode()))

    def load_scene(self, path: str):
        self._ck(self.L.tfb_scene_load(self.h, path.encode()))

    # -- block streaming between the voxel pool and host memory (tfb_stream_*, include/tfusion_b200.h) --
    def stream_out(self, max_blocks: int = 0) -> int:
        n = C.c_int(0)
        self._ck(self.L.tfb_stream_out(self.h, C.c_int(max_blocks), C.byref(n)))
        return n.value

    def stream_in(self, pose_w2c=None, all_blocks: bool = False):
        """returns (blocks restored, blocks left in the host store)"""
        n, left = C.c_int(0), C.c_int(0)
        p = None
        if pose_w2c is not None:
            m = np.ascontiguousarray(pose_w2c, dtype=np.float32)
            p = m.ctypes.data_as(C.c_void_p)
        self._ck(self.L.tfb_stream_in(self.h, p, C.c_int(1 if all_blocks else 0), C.byref(n), C.byref(left)))
        return n.value, left.value

    def stream_stats(self):
        """(blocks in the pool, blocks in the host store)"""
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._ck(self.L.tfb_stream_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- frames ----------------------------------------------------------------------------------------
    def process_frame(self, depth) -> bool:
        """depth: host u16 [rows, cols]; a PinnedArray's .array makes the upload asynchronous."""
        d = depth if (isinstance(depth, np.ndarray) and depth.dtype == np.uint16 and depth.flags["C_CONTIGUOUS"]) \
            else np.ascontiguousarray(depth, dtype=np.uint16)
        ok = C.c_int(0)
        self._ck(self.L.tfb_process_frame(self.h, _np_ptr(d), C.c_size_t(d.strides[0]), C.byref(ok)))
        return bool(ok.value)

    def process_frame_device(self, dev_ptr) -> bool:
        ok = C.c_int(0)
        p = dev_ptr.ptr if isinstance(dev_ptr, DevBuf) else C.c_void_p(dev_ptr)
        self._ck(self.L.tfb_process_frame_device(self.h, p, C.byref(ok)))
        return bool(ok.value)

    # -- sharded scene: the frame in three stages, a cross-GPU barrier on the stream between them -------
    def stream(self) -> int:
        return int(self.L.tfb_stream(self.h) or 0)

    def shard_local_ptrs(self) -> ShardPtrs:
        p = ShardPtrs()
        self._ck(self.L.tfb_shard_local_ptrs(self.h, C.byref(p)))
        return p

    def shard_attach(self, rank: int, ptrs: ShardPtrs):
        self._ck(self.L.tfb_shard_attach(self.h, C.c_int(rank), C.byref(ptrs)))

    def frame_begin(self, dev_ptr=None):
        """dev_ptr None: the frame rank 0 pushed into this context's frame buffer (shard_push_frame)"""
        p = None if dev_ptr is None else (dev_ptr.ptr if isinstance(dev_ptr, DevBuf) else C.c_void_p(dev_ptr))
        self._ck(self.L.tfb_frame_begin(self.h, p))

    def shard_push_frame(self, dev_ptr):
        p = dev_ptr.ptr if isinstance(dev_ptr, DevBuf) else C.c_void_p(dev_ptr)
        self._ck(self.L.tfb_shard_push_frame(self.h, p))

    def process_frame_sharded(self, dev_ptr=None) -> bool:
        """collective: the whole sharded frame; the rank that holds the frame passes it, the others None"""
        ok = C.c_int(0)
        p = None if dev_ptr is None else (dev_ptr.ptr if isinstance(dev_ptr, DevBuf) else C.c_void_p(dev_ptr))
        self._ck(self.L.tfb_process_frame_sharded(self.h, p, C.byref(ok)))
        return bool(ok.value)

    def shard_barrier(self):
        self._ck(self.L.tfb_shard_barrier(self.h))

    def frame_raycast(self):
        self._ck(self.L.tfb_frame_raycast(self.h))

    def frame_end(self) -> bool:
        ok = C.c_int(0)
        self._ck(self.L.tfb_frame_end(self.h, C.byref(ok)))
        return bool(ok.value)

    def num_poses(self) -> int:
        return int(self.L.tfb_num_poses(self.h))

    def pose(self, idx: int = -1):
        out = np.empty(16, np.float32)
        self._ck(self.L.tfb_get_pose(self.h, C.c_int(idx), _np_ptr(out)))
        return out.reshape(4, 4)

    # -- inspection ------------------------------------------------------------------------------------
    def counters(self) -> dict:
        c = np.zeros(8, np.int64)
        self._ck(self.L.tfb_get_counters(self.h, _np_ptr(c)))
        keys = ["n_visible", "last_free_block", "last_free_excess", "n_new_frame", "frame_counter", "resets",
                "n_raycast_extras", "n_allocated"]
        return dict(zip(keys, (int(v) for v in c)))

    def voxel_updates(self) -> int:
        return int(self.L.tfb_voxel_updates_last(self.h))

    def voxel_updates_total(self) -> int:
        """summed over every integration finished so far; never waits for a deferred tail"""
        return int(self.L.tfb_voxel_updates_total(self.h))

    def kernel_launches(self) -> int:
        return int(self.L.tfb_kernel_launches(self.h))

    def table(self):
        n = int(self.L.tfb_total_entries(self.h))
        t = np.empty(n, HASH_DTYPE)
        self._ck(self.L.tfb_export_table(self.h, _np_ptr(t)))
        return t

    def vis_type(self):
        n = int(self.L.tfb_total_entries(self.h))
        t = np.empty(n, np.uint8)
        self._ck(self.L.tfb_export_vis_type(self.h, _np_ptr(t)))
        return t

    def visible_ids(self):
        cap = int(self.L.tfb_total_entries(self.h))
        ids = np.empty(cap, np.int32)
        n = C.c_int(0)
        self._ck(self.L.tfb_export_visible_ids(self.h, _np_ptr(ids), C.c_int(cap), C.byref(n)))
        return ids[: n.value].copy()

    def block(self, ptr: int):
        b = np.empty(512, VOXEL_DTYPE)
        self._ck(self.L.tfb_export_block(self.h, C.c_int(ptr), _np_ptr(b)))
        return b

    def blocks_by_pos(self) -> dict:
        t = self.table()
        out = {}
        for e in t[t["ptr"] >= 0]:
            out[tuple(int(v) for v in e["pos"])] = self.block(int(e["ptr"]))
        return out

    def minmax(self):
        m = np.empty((self.rows // 8, self.cols // 8, 2), np.float32)
        self._ck(self.L.tfb_export_minmax(self.h, _np_ptr(m)))
        return m

    def raycast_result(self):
        m = np.empty((self.rows, self.cols, 4), np.float32)
        self._ck(self.L.tfb_export_raycast(self.h, _np_ptr(m)))
        return m

    def dists(self):
        m = np.empty((self.rows, self.cols), np.float32)
        self._ck(self.L.tfb_export_dists(self.h, _np_ptr(m)))
        return m

    def level(self, which: int, level: int):
        w, h = self.cols >> level, self.rows >> level
        out = np.empty((h, w), np.uint16) if which == 0 else np.empty((h, w, 4), np.float32)
        self._ck(self.L.tfb_export_level(self.h, C.c_int(which), C.c_int(level), _np_ptr(out)))
        return out

    def set_level(self, which: int, level: int, arr):
        a = np.ascontiguousarray(arr, dtype=np.uint16 if which == 0 else np.float32)
        self._ck(self.L.tfb_import_level(self.h, C.c_int(which), C.c_int(level), _np_ptr(a)))

    def icp_valid_list(self) -> np.ndarray:
        """level-0 pixels that hold a vertex, ascending, as the last tracked frame's ICP shared them out"""
        out = np.empty(self.cols * self.rows, np.int32)
        n = C.c_int(0)
        self._ck(self.L.tfb_export_icp_valid_list(self.h, _np_ptr(out), C.c_int(out.size), C.byref(n)))
        return out[:n.value].copy()

    def level_ptr(self, which: int, level: int) -> int:
        return int(self.L.tfb_level_ptr(self.h, C.c_int(which), C.c_int(level)))

    def mark(self, slot: int):
        self._ck(self.L.tfb_mark(self.h, C.c_int(slot)))

    def elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        self._ck(self.L.tfb_elapsed_ms(self.h, C.c_int(a), C.c_int(b), C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        self._ck(self.L.tfb_flush_l2(self.h))

    def ktiming(self, on: bool = True, reset: bool = True):
        self._ck(self.L.tfb_ktiming_enable(self.h, C.c_int(1 if on else 0)))
        if reset:
            self._ck(self.L.tfb_ktiming_reset(self.h))

    def kernel_times(self) -> dict:
        """{kernel name: (total_ms, launches)} accumulated while ktiming was on"""
        out = {}
        for i in range(int(self.L.tfb_ktiming_count())):
            ms, n = C.c_double(0), C.c_longlong(0)
            self._ck(self.L.tfb_ktiming_get(self.h, C.c_int(i), C.byref(ms), C.byref(n)))
            if n.value:
                out[self.L.tfb_ktiming_name(C.c_int(i)).decode()] = (float(ms.value), int(n.value))
        return out

    def timing(self, on: bool = True):
        self._ck(self.L.tfb_timing_enable(self.h, C.c_int(1 if on else 0)))

    def stage_ms(self) -> dict:
        t = np.zeros(9, np.float32)
        self._ck(self.L.tfb_timing_last_ms(self.h, _np_ptr(t)))
        return dict(zip(STAGES, (float(v) for v in t)))


def ipc_export(ptr: int) -> bytes:
    """cudaIpcGetMemHandle of a device allocation of this process (64 opaque bytes to send to the other ranks)"""
    buf = C.create_string_buffer(64)
    if lib().tfb_ipc_export(C.c_void_p(ptr), buf) != 0:
        raise TfbError("cudaIpcGetMemHandle failed")
    return buf.raw


def ipc_open(handle: bytes) -> int:
    """cudaIpcOpenMemHandle: a pointer, valid in this process, to another rank's allocation (peer access enabled)"""
    out = C.c_void_p()
    if lib().tfb_ipc_open(C.create_string_buffer(handle, 64), C.byref(out)) != 0:
        raise TfbError("cudaIpcOpenMemHandle failed (are the GPUs peer-accessible?)")
    return int(out.value)


def allocated_set(table) -> set:
    t = table[table["ptr"] >= -1]
    return {tuple(int(v) for v in e) for e in t["pos"]}


def visible_set(table, ids) -> set:
    """visible list as a set of allocated block coordinates."""
    e = table[ids]
    e = e[e["ptr"] >= -1]
    return {tuple(int(v) for v in p) for p in e["pos"]}
