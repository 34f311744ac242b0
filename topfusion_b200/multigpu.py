"""One TopFu whose voxel-block scene is sharded over the GPUs of one node.

One process per GPU (torchrun); `torch.distributed` is the plumbing (NCCL on the GPUs, gloo in the CPU tests):

  * the depth frame is broadcast from rank 0 (0.6 MB at 640x480);
  * every rank runs preprocess + ICP + allocation redundantly — they are deterministic, so the hash INDEX is a replica
    on every rank and the admission order of new blocks is the one of a single GPU;
  * a block's 2 KB voxel payload lives only in the pool of its owner (`owner_rank`), which integrates it: no exchange;
  * after a barrier every rank casts its share of the image rows and reads the voxels of foreign blocks straight out
    of the owner's table + pool over NVLink peer memory (CUDA IPC pointers), storing finished rows and visibility marks
    into every rank's buffers from inside the same kernel; after a second barrier all replicas hold the same raycast
    image and the same visible set, and derive the same model maps for the next frame's ICP.

The result is the single-GPU result bit for bit (tests/test_gpu_sharding.py); what shards is the integration work,
the raycast work and the voxel memory (N x 180 GB).  ICP stays replicated: its payload is 27 floats.

There is no CPU fallback: the CUDA engine needs libtfusion_b200.so and one GPU per rank.  The CPU tests
(tests/test_sharding_gloo.py) drive the same host logic with an engine built on the oracle.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def owner_rank(bx: int, by: int, bz: int, count: int) -> int:
    """payload owner of block (bx,by,bz): tfb::owner_rank (csrc/tfb_common.cuh), uint32 arithmetic"""
    if count <= 1:
        return 0
    m = 0xFFFFFFFF
    h = ((bx * 0x9E3779B1) & m) ^ ((by * 0x85EBCA77) & m) ^ ((bz * 0xC2B2AE3D) & m)
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & m
    h ^= h >> 12
    return h % count


def rows_of_rank(rows: int, rank: int, count: int, strip: int = 8):
    """image rows rank `rank` casts: every count-th strip of 8 rows (k_raycast_sharded)"""
    out = []
    for s in range((rows + strip - 1) // strip):
        if s % count == rank:
            out.extend(range(s * strip, min(rows, (s + 1) * strip)))
    return out


class CudaEngine:
    """this rank's tfb context on its GPU, attached to every other rank's buffers through CUDA IPC"""

    def __init__(self, rank: int, world: int, device_index: int, dist, group=None, plumbing: str = "p2p", **params):
        """plumbing: "p2p" — frame push and barriers are this library's own kernels over peer memory (tfb_shard_push_frame /
        tfb_shard_barrier); "nccl" — torch.distributed broadcast and one-element all-reduces on the same stream."""
        self.plumbing = plumbing if world > 1 else "none"
        import torch
        from . import capi
        self.capi, self.torch = capi, torch
        torch.cuda.set_device(device_index)
        capi.lib().tfb_set_device(device_index)
        self.stream = torch.cuda.Stream(device=device_index)
        self.ctx = capi.Context(stream=self.stream.cuda_stream, shard_rank=rank, shard_count=world, **params)
        self.rank, self.world = rank, world
        self._opened = []
        if world > 1:
            mine = self.ctx.shard_local_ptrs()
            handles = [capi.ipc_export(getattr(mine, f)) for f, _ in capi.ShardPtrs._fields_]
            everyone = [None] * world
            dist.all_gather_object(everyone, handles, group=group)
            for r in range(world):
                if r == rank:
                    continue
                p = capi.ShardPtrs()
                for (f, _), h in zip(capi.ShardPtrs._fields_, everyone[r]):
                    ptr = capi.ipc_open(h)
                    self._opened.append(ptr)
                    setattr(p, f, ptr)
                self.ctx.shard_attach(r, p)

    def begin(self, frame):
        # device tensor, rows x cols u16 (viewed as int16 by torch); with the p2p plumbing rank 0 has already pushed it into
        # every rank's own frame buffer
        self.ctx.frame_begin(None if self.plumbing == "p2p" else frame.data_ptr())

    def process_collective(self, frame_or_none) -> bool:
        """p2p plumbing: push + barriers + the software-pipelined frame in one library call (tfb_process_frame_sharded)"""
        return self.ctx.process_frame_sharded(None if frame_or_none is None else frame_or_none.data_ptr())

    def push_frame(self, frame):
        self.ctx.shard_push_frame(frame.data_ptr())

    def barrier(self):
        self.ctx.shard_barrier()

    def raycast(self):
        self.ctx.frame_raycast()

    def end(self) -> bool:
        return self.ctx.frame_end()

    def voxel_updates(self) -> int:
        return self.ctx.voxel_updates()

    def pose(self):
        return self.ctx.pose()

    def close(self):
        if self.ctx is not None:
            self.ctx.sync()
            for p in self._opened:
                self.capi.lib().tfb_ipc_close(self.capi.C.c_void_p(p))
            self._opened = []
            self.ctx.close()
            self.ctx = None


class ShardedTopFu:
    """rank-local handle of the sharded reconstruction: process_frame() has TopFu::operator()'s meaning on every rank.

    engine: anything with begin(frame) / raycast() / end() -> bool / voxel_updates(); CudaEngine on the GPUs.
    frame_buffer: a tensor the frame is broadcast into (device tensor for NCCL, CPU tensor for gloo)."""

    def __init__(self, engine, dist, rank: int, world: int, frame_buffer, group=None):
        import torch
        self.torch, self.dist, self.group = torch, dist, group
        self.engine, self.rank, self.world = engine, rank, world
        self.frame = frame_buffer
        self._flag = torch.zeros(1, dtype=torch.int32, device=frame_buffer.device)
        self.frames_done = 0
        # TFB_SHARD_PIPELINE=0: the three stages as separate calls with a barrier in between (what the NCCL plumbing needs)
        self.pipelined = os.environ.get("TFB_SHARD_PIPELINE", "1") != "0"

    def _own_plumbing(self) -> bool:
        return getattr(self.engine, "plumbing", "") == "p2p"

    def barrier(self):
        """cross-rank barrier ON THE STREAM (no host synchronisation): the engine's own flag barrier over peer memory, or a
        one-element all-reduce"""
        if self.world <= 1:
            return
        if self._own_plumbing():
            self.engine.barrier()
        else:
            self.dist.all_reduce(self._flag, group=self.group)

    def distribute(self, frame_src=None):
        """rank 0 copies its frame into the buffer (H2D when it is a pinned host tensor), everyone receives it"""
        if self.rank == 0 and frame_src is not None:
            self.frame.copy_(frame_src, non_blocking=True)
        if self.world <= 1:
            return
        if self._own_plumbing():
            if self.rank == 0:
                self.engine.push_frame(self.frame)   # one kernel: NVLink stores into every rank's frame buffer
            self.barrier()                            # the frame has arrived everywhere
        else:
            self.dist.broadcast(self.frame.view(self.torch.uint8), src=0, group=self.group)   # bytes: every backend moves u8

    def process_frame(self, frame_src=None) -> bool:
        if self.world > 1 and self._own_plumbing() and self.pipelined and hasattr(self.engine, "process_collective"):
            if self.rank == 0 and frame_src is not None:
                self.frame.copy_(frame_src, non_blocking=True)
            ok = self.engine.process_collective(self.frame if self.rank == 0 else None)
            self.frames_done += 1
            return ok
        self.distribute(frame_src)
        self.engine.begin(self.frame)
        self.barrier()        # every owner has integrated: voxels are final
        self.engine.raycast()
        self.barrier()        # every rank's rows and marks have arrived
        ok = self.engine.end()
        self.frames_done += 1
        return ok

    def total(self, value: float, op: str = "sum") -> float:
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self._flag.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op.upper()), group=self.group)
        return float(t.item())


# ---------------------------------------------------------------------------------------------------------
# voxel-updates/s on the large scene (BASELINE.json configs[3]): the part of the path that shards with no exchange at all
# ---------------------------------------------------------------------------------------------------------
LARGE_SCENE = "S3 large scene (2x1x2 m room shell + 20 boxes), 640x480, 2 mm voxels, mu 16 mm, ground-truth poses"


def integrate_scaling_leg(rank: int, world: int, reduce_max=None, reduce_sum=None, frames: int = 24, warm: int = 4,
                          cols: int = 640, rows: int = 480):
    """every rank allocates the replicated index and integrates the blocks it owns from the same frames; no collective on
    the data path.  Returns aggregate voxel-updates/s = updates of all ranks / slowest rank's integration time (k_integrate
    plus, on a sharded scene, the k_owned_list pass in front of it).  cols x rows: the sensor resolution — 640x480 exposes
    ~50 k visible blocks per frame, 1280x720 ~4x that, which is what keeps 8 GPUs out of the launch-latency regime."""
    import ctypes as C
    from . import capi, synth
    n_seq = 12 if cols <= 640 else 6      # the analytic renderer costs ~4 s per 1280x720 frame of this scene
    cache = os.path.join("/tmp", "tfb_s3_12.npz" if (cols, rows) == (640, 480) else f"tfb_s3_{n_seq}_{cols}x{rows}.npz")
    depth = poses = None
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            depth, poses = z["depth"], z["poses"]
        except Exception:      # another rank is writing it right now
            depth = None
    intr = synth.intrinsics_for(cols, rows)
    if depth is None:
        depth, poses, _ = synth.sequence("S3", n_seq, cols, rows)
        if rank == 0:
            tmp = f"{cache}.{os.getpid()}.tmp.npz"
            try:
                np.savez(tmp, depth=depth, poses=poses)
                os.replace(tmp, cache)
            except OSError:
                pass
    ctx = capi.Context(cols=cols, rows=rows, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3],
                       voxel_size=0.002, mu=0.016, num_blocks=1 << (19 if cols <= 640 else 20), num_buckets=1 << 22, excess_size=1 << 18,
                       depth_cutoff_mm=4000, shard_rank=rank, shard_count=world)
    dev = [ctx.upload(depth[i]) for i in range(n_seq)]
    dists = capi.DevBuf(depth.shape[1] * depth.shape[2] * 4)
    order, i, step = [], 0, 1
    while len(order) < warm + frames:
        order.append(i)
        if i + step >= n_seq or i + step < 0:
            step = -step
        i += step
    upd = 0
    for t, fi in enumerate(order):
        if t == warm:
            ctx.ktiming(True)
        w2c = np.ascontiguousarray(np.linalg.inv(poses[fi]), dtype=np.float32)
        ctx._ck(ctx.L.tfb_compute_dists(ctx.h, dev[fi].ptr, dists.ptr, C.c_int(depth.shape[2]), C.c_int(depth.shape[1])))
        ctx._ck(ctx.L.tfb_allocate_scene_from_depth(ctx.h, w2c.ctypes.data_as(C.c_void_p), dists.ptr))
        ctx.flush_l2()
        ctx._ck(ctx.L.tfb_integrate_into_scene(ctx.h, w2c.ctypes.data_as(C.c_void_p), dists.ptr))
        if t >= warm:
            upd += ctx.voxel_updates()
    kt = ctx.kernel_times()
    ms, launches = kt["k_integrate"]
    ms_owned = kt.get("k_owned_list", (0.0, 0))[0]
    ms += ms_owned          # the compaction pass belongs to the integration stage of a sharded scene
    ctx.close()
    ms_max = reduce_max(ms) if reduce_max else ms
    upd_all = reduce_sum(upd) if reduce_sum else upd
    peak = 6542.7
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    per_s = upd_all / (ms_max / 1000.0)
    return {"workload": LARGE_SCENE.replace("640x480", f"{cols}x{rows}"), "value": per_s, "unit": "voxel-updates/s", "frames": frames,
            "visible_blocks_per_frame_all_ranks": upd_all / 512.0 / frames, "k_integrate_us_slowest_rank": 1000.0 * ms_max / launches,
            "algorithmic_gbs_all_ranks": per_s * 8.04 / 1e9, "frac_of_measured_hbm_peak_per_gpu": per_s * 8.04 / 1e9 / world / peak,
            "k_owned_list_us_this_rank": 1000.0 * ms_owned / max(launches, 1),
            "l2": "flushed before every integration"}


# ---------------------------------------------------------------------------------------------------------
# bench.py --gpus N (launched by torchrun, one rank per GPU)
# ---------------------------------------------------------------------------------------------------------
def bench_main(args, rank, world, local_rank, orbit_frames, ClockSampler, workload):
    import torch
    import torch.distributed as dist
    from . import capi

    # stdout carries exactly one JSON line: NCCL and the launcher print there too, so park it on stderr until the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    W, K = args.warmup, args.steps
    frames, gt = orbit_frames(W + K)
    rows, cols = frames.shape[1], frames.shape[2]
    dev = torch.device("cuda", local_rank)

    def run(leg: str):
        eng = CudaEngine(rank, world, local_rank, dist, plumbing=os.environ.get("TFB_PLUMBING", "p2p"), corrected_mode=args.mode)
        with torch.cuda.stream(eng.stream):
            buf = torch.empty((rows, cols), dtype=torch.int16, device=dev)
            st = ShardedTopFu(eng, dist, rank, world, buf)
            if leg == "resident":
                src = [torch.from_numpy(frames[i].view(np.int16)).to(dev) for i in range(W + K)] if rank == 0 else None
            else:
                src = [torch.from_numpy(frames[i].view(np.int16)).pin_memory() for i in range(W + K)] if rank == 0 else None
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            vox = 0
            oks = 0
            l0 = 0
            dist.barrier()
            torch.cuda.synchronize()
            for i in range(W + K):
                eng.ctx.flush_l2()
                if i == W:
                    l0 = eng.ctx.kernel_launches()
                if i >= W:
                    ev[i - W][0].record(eng.stream)
                ok = st.process_frame(src[i] if rank == 0 else None)
                if leg == "e2e":
                    _ = eng.pose()
                if i == W:
                    vox = -eng.ctx.voxel_updates_total()
                if i >= W:
                    ev[i - W][1].record(eng.stream)
                    oks += int(ok)
            vox += eng.ctx.voxel_updates_total()   # integrations finished by the timed calls (host counter, does not wait)
            dist.barrier()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev)
            launches = eng.ctx.kernel_launches() - l0
            ms_max = st.total(ms, "max")
            vox_all = st.total(vox, "sum")
            pose_err = float(np.abs(eng.pose()[:3, 3] - gt[W + K - 1][:3, 3]).max())
            n_alloc = st.total(eng.ctx.counters()["n_allocated"], "sum")
        eng.close()
        return ms_max, vox_all, oks, launches, pose_err, n_alloc

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, vox, oks, launches, pose_err, n_alloc = run("resident")
    clocks = sampler.result()
    e_ms, _, _, _, _, _ = run("e2e")

    def red(op):
        def f(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=op)
            return float(t.item())
        return f
    large = integrate_scaling_leg(rank, world, red(dist.ReduceOp.MAX), red(dist.ReduceOp.SUM))
    large_720p = integrate_scaling_leg(rank, world, red(dist.ReduceOp.MAX), red(dist.ReduceOp.SUM), frames=16, cols=1280, rows=720)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        line = {
            "metric": "frames/sec (ICP+integrate+raycast, 640x480)", "value": K / (ms / 1000.0), "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "mode": "corrected" if args.mode else "reference", "voxel_size_m": 0.005,
                       "parallelism": f"scene sharded over {world} GPUs by block-coordinate hash: index replicated, "
                                      "payload + integration + raycast rows partitioned, peer-memory raycast; ICP replicated; "
                                      "frame push and barriers: " + os.environ.get("TFB_PLUMBING", "p2p"),
                       "l2": "256 MB scratch overwritten between timed steps (L2 flushed)",
                       "final_pose_err_m": pose_err, "frames_tracked": oks, "blocks_allocated_all_ranks": n_alloc},
            "e2e": {"value": K / (e_ms / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": rows * cols * 2,
                    "d2h_bytes_per_step": 468, "ms_per_step": e_ms / K},
            "gpu_launches": launches, "clocks": clocks,
            "voxel_updates_per_s": vox / (ms / 1000.0),
            "voxel_updates_large_scene": large,
            "voxel_updates_large_scene_1280x720": large_720p,
            "roofline": {"bound": "hbm", "kernel": "k_integrate", "achieved": large["algorithmic_gbs_all_ranks"] / world,
                         "peak": large["algorithmic_gbs_all_ranks"] / world / max(large["frac_of_measured_hbm_peak_per_gpu"], 1e-12),
                         "unit": "GB/s", "frac": large["frac_of_measured_hbm_peak_per_gpu"], "traffic": None,
                         "note": "per GPU, on the large-scene integration leg (the stage that shards without exchange); "
                                 "the N=1 line carries the frame's dominant kernel"},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
