#!/usr/bin/env python
"""BASELINE.json configs[2]: synthetic 1280x720 room-scale walk-through (S2), 4 mm voxels, 2 M-block pool, 2^22-bucket hash,
on one B200.  Same timing protocol as bench.py (CUDA events around every call, L2 flushed between calls); prints one JSON line.

    python tools/bench_config3.py [--frames 40] [--warmup 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    import bench
    from topfusion_b200 import capi, synth
    n = a.frames + a.warmup
    cache = f"/tmp/tfb_s2_{n}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        depth, poses, intr = z["depth"], z["poses"], tuple(z["intr"])
    else:
        depth, poses, intr = synth.sequence("S2", n)
        np.savez(cache, depth=depth, poses=poses, intr=np.array(intr))
    kw = dict(cols=1280, rows=720, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], voxel_size=0.004, mu=0.016,
              num_blocks=1 << 21, num_buckets=1 << 22, excess_size=1 << 19, depth_cutoff_mm=4000, view_frustum_max=4.0,
              icp_truncate_depth_dist=4.0, corrected_mode=1)
    ctx = capi.Context(**kw)
    bufs = [ctx.upload(depth[i]) for i in range(n)]
    ms, vox, oks = bench.timed_loop(ctx, lambda i: ctx.process_frame_device(bufs[i]), a.warmup, a.frames, flush=True)
    err = float(np.abs(ctx.pose()[:3, 3] - poses[n - 1][:3, 3]).max())
    cnt = ctx.counters()
    ctx.close()
    ctx = capi.Context(**kw)
    for i in range(a.warmup):
        ctx.process_frame_device(bufs[i])
    ctx.ktiming(True)
    k = min(a.frames, 20)
    for i in range(a.warmup, a.warmup + k):
        ctx.flush_l2()
        ctx.process_frame_device(bufs[i])
    kt = {name: round(1000.0 * v[0] / v[1], 2) for name, v in sorted(ctx.kernel_times().items(), key=lambda kv: -kv[1][0])}
    ctx.close()
    print(json.dumps({"workload": "S2 synthetic 1280x720 room walk-through, 4 mm voxels, mu 16 mm, 2M-block pool, 2^22 buckets, ICP {10,5,4}",
                      "value": a.frames / (ms / 1000.0), "unit": "frames/s", "ms_per_frame": ms / a.frames, "frames": a.frames,
                      "frames_tracked": oks, "final_pose_err_m": err, "visible_blocks_end": cnt["n_visible"],
                      "blocks_allocated": cnt["n_allocated"], "voxel_updates_per_s": vox / (ms / 1000.0),
                      "us_per_launch": kt, "l2": "flushed between calls"}))


if __name__ == "__main__":
    main()
