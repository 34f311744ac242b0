#!/usr/bin/env python
"""ONE GPU: the sharded frame of N emulated ranks (N contexts on one stream, attached by plain device pointers — the
staged entry points, stream order as the barrier), with the library's per-kernel event timing.

Purpose: k_raycast_sharded / k_gather_foreign WITHOUT NVLink — every "peer" read and store is local — so the difference
to the real 2-GPU run (tools/shard_kernel_times.py) is the interconnect and the difference to k_raycast on one context is
the code shape.  Also the command to put under `ncu --set full -k regex:"k_raycast_sharded|k_gather_foreign"`; ncu must
never be wrapped around a multi-rank run.

    python tools/profile_sharded_emulated.py [--ranks 2] [--frames 24]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=2)
    ap.add_argument("--frames", type=int, default=24)
    args = ap.parse_args()
    from topfusion_b200 import capi, synth
    n, warm = args.frames, 6
    depth, _, _ = synth.sequence("S1", n + warm)
    c0 = capi.Context(shard_rank=0, shard_count=args.ranks, corrected_mode=1)
    ctxs = [c0] + [capi.Context(stream=c0.stream(), shard_rank=r, shard_count=args.ranks, corrected_mode=1)
                   for r in range(1, args.ranks)]
    single = capi.Context(corrected_mode=1, defer_tail=0)
    ptrs = [c.shard_local_ptrs() for c in ctxs]
    for c in ctxs:
        for r, p in enumerate(ptrs):
            c.shard_attach(r, p)
    for i in range(n + warm):
        if i == warm:
            for c in ctxs + [single]:
                c.sync()
                c.ktiming(True)
        buf = c0.upload(depth[i], "frame")
        for c in ctxs:
            c.frame_begin(buf)
        for c in ctxs:
            c.frame_raycast()
        oks = [c.frame_end() for c in ctxs]
        ok1 = single.process_frame(depth[i])
        assert all(o == ok1 for o in oks)
    out = {"frames": n, "ranks": args.ranks,
           "emulated_rank_us_per_frame": [{k: round(1000.0 * ms / n, 2) for k, (ms, _) in c.kernel_times().items()} for c in ctxs],
           "single_context_us_per_frame": {k: round(1000.0 * ms / n, 2) for k, (ms, _) in single.kernel_times().items()}}
    print(json.dumps(out))
    for c in reversed(ctxs):
        c.close()
    single.close()


if __name__ == "__main__":
    main()
