#!/usr/bin/env python
"""Integration / raycast microbench (BASELINE.json configs[4]): fixed ground-truth poses, the orbit
cycled for N frames, voxel size and truncation band swept.  Per kernel: CUDA-event time per launch, ALGORITHMIC bytes
 and the fraction of the measured HBM peak; voxel-updates/s = 512 x allocated visible blocks / k_integrate time.

    python tools/microbench.py --voxel-mm 2 3 5 --mu-voxels 4 8 --frames 60 [--seq S1|S2|S3] [--no-flush]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(capi, depth, poses, intr, voxel, mu, frames, warm, flush, depth_cutoff_mm, ieee=0, shard=None):
    rows, cols = depth.shape[1:]
    ctx = capi.Context(cols=cols, rows=rows, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], voxel_size=voxel, mu=mu,
                       num_blocks=1 << 20, num_buckets=1 << 22, excess_size=1 << 19, depth_cutoff_mm=depth_cutoff_mm, ieee_arith=ieee,
                       view_frustum_max=max(3.0, depth_cutoff_mm / 1000.0 + 2 * mu),
                       **({"shard_count": shard[0], "shard_rank": shard[1]} if shard else {}))
    L = ctx.L
    n = depth.shape[0]
    dev = [ctx.upload(depth[i]) for i in range(n)]
    dists = capi.DevBuf(rows * cols * 4)
    pts, nrm = capi.DevBuf(rows * cols * 16), capi.DevBuf(rows * cols * 16)
    nblk = []
    order = []
    i, step = 0, 1
    while len(order) < warm + frames:
        order.append(i)
        if i + step >= n or i + step < 0:
            step = -step
        i += step
    for t, fi in enumerate(order):
        if t == warm:
            ctx.ktiming(True)
        c2w = np.ascontiguousarray(poses[fi], dtype=np.float32)
        w2c = np.ascontiguousarray(np.linalg.inv(poses[fi]), dtype=np.float32)
        if flush == "frame":   # cold at the start of the frame: the voxels come from HBM, what the frame's own stages produce stays in L2
            ctx.flush_l2()
        ctx._ck(L.tfb_compute_dists(ctx.h, dev[fi].ptr, dists.ptr, C.c_int(cols), C.c_int(rows)))
        ctx._ck(L.tfb_allocate_scene_from_depth(ctx.h, w2c.ctypes.data_as(C.c_void_p), dists.ptr))
        if flush is True:
            ctx.flush_l2()
        ctx._ck(L.tfb_integrate_into_scene(ctx.h, w2c.ctypes.data_as(C.c_void_p), dists.ptr))
        if t >= warm:
            nblk.append(ctx.voxel_updates() / 512.0)
        if not shard:   # one rank of a sharded scene alone: allocation + integration only (the raycast reads the other ranks)
            ctx._ck(L.tfb_create_expected_depths(ctx.h, w2c.ctypes.data_as(C.c_void_p)))
            if flush is True:
                ctx.flush_l2()
            ctx._ck(L.tfb_create_icp_maps(ctx.h, c2w.ctypes.data_as(C.c_void_p), pts.ptr, nrm.ptr))
        ctx.sync()
    kt = ctx.kernel_times()
    cnt = ctx.counters()
    ctx.close()
    nb = float(np.mean(nblk))
    p0 = rows * cols
    alg = {"k_integrate": nb * 4116.0 + 4.0 * p0, "k_raycast": p0 * 16.0 + p0 / 64.0 * 8.0 + nb * 2064.0, "k_icp_maps": p0 * 48.0,
           "k_mark": 4.0 * p0 + 16.0 * nb, "k_expected_depths": 36.0 * nb}
    out = {"voxel_mm": voxel * 1000, "mu_mm": mu * 1000, "cols": cols, "rows": rows, "visible_blocks_avg": nb,
           "allocated": cnt["n_allocated"], "l2_flushed": bool(flush), "ieee_arith": int(ieee), "kernels": {}}
    pk = peak()
    for k, (ms, launches) in kt.items():
        us = 1000.0 * ms / launches
        e = {"us_per_launch": us}
        if k in alg:
            e["algorithmic_bytes"] = alg[k]
            e["gbs"] = alg[k] / (us * 1e-6) / 1e9
            e["frac_of_measured_hbm_peak"] = e["gbs"] / pk
        out["kernels"][k] = e
    out["voxel_updates_per_s"] = nb * 512.0 / (out["kernels"]["k_integrate"]["us_per_launch"] * 1e-6)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxel-mm", type=float, nargs="+", default=[5.0])
    ap.add_argument("--mu-voxels", type=float, nargs="+", default=[4.0])
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--seq", default="S1")
    ap.add_argument("--seq-frames", type=int, default=40)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--flush-frame", action="store_true", help="flush L2 once per frame (before the allocation stage) instead of in front of every timed kernel")
    ap.add_argument("--out", default=None)
    ap.add_argument("--ieee", type=int, default=0, help="1: IEEE integration arithmetic (tfb_params.ieee_arith)")
    ap.add_argument("--shard", type=int, nargs=2, default=None, metavar=("COUNT", "RANK"),
                    help="time ONE rank of a scene sharded COUNT ways on this GPU (allocation + integration only)")
    a = ap.parse_args()
    from topfusion_b200 import capi, synth
    depth, poses, intr = synth.sequence(a.seq, a.seq_frames)
    cutoff = 2047 if a.seq in ("S0", "S1") else 4000
    res = []
    for v in a.voxel_mm:
        for m in a.mu_voxels:
            r = run(capi, depth, poses, intr, v / 1000.0, m * v / 1000.0, a.frames, a.warmup, ("frame" if a.flush_frame else not a.no_flush), cutoff, a.ieee, a.shard)
            res.append(r)
            ki = r["kernels"]["k_integrate"]
            kr = r["kernels"].get("k_raycast", {"us_per_launch": 0.0, "gbs": 0.0, "frac_of_measured_hbm_peak": 0.0})
            print(f"{a.seq} voxel {v} mm mu {m * v} mm: {r['visible_blocks_avg']:.0f} blocks | integrate {ki['us_per_launch']:.1f} us "
                  f"{ki['gbs']:.0f} GB/s ({100 * ki['frac_of_measured_hbm_peak']:.1f} %) {r['voxel_updates_per_s'] / 1e9:.1f} G upd/s | "
                  f"raycast {kr['us_per_launch']:.1f} us {kr['gbs']:.0f} GB/s ({100 * kr['frac_of_measured_hbm_peak']:.1f} %)", flush=True)
    if a.out:
        json.dump({"seq": a.seq, "peak_gbs": peak(), "results": res}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
