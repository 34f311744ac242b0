"""GPU box: frames/s of the headline loop (S1 orbit, 640x480, 5 mm, corrected mode, frames resident in HBM, L2 flushed between
steps), then per-kernel times of a second pass.  TFB_LIB_PATH selects a variant build (A/B in one gpurun call).
usage: python tools/fps_quick.py [n_timed] [repeats]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from topfusion_b200 import capi, synth

n_timed = int(sys.argv[1]) if len(sys.argv) > 1 else 90
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W = 10
depth, _, _ = synth.sequence("S1", 100)
res = []
for r in range(reps):
    ctx = capi.Context(corrected_mode=1)
    bufs = [ctx.upload(depth[i]) for i in range(W + n_timed)]
    for i in range(W):
        ctx.process_frame_device(bufs[i])
    ctx.sync()
    ms = 0.0
    for i in range(W, W + n_timed):
        ctx.flush_l2()
        ctx.mark(0)
        ok = ctx.process_frame_device(bufs[i])
        ctx.mark(1)
        ms += ctx.elapsed_ms(0, 1)
        assert ok
    res.append(n_timed / (ms / 1000.0))
    if r == reps - 1:
        ctx.ktiming(True)
        for i in range(W, W + 30):
            ctx.flush_l2()
            ctx.process_frame_device(bufs[i])
        kt = ctx.kernel_times()
        print({k: round(1000.0 * v[0] / v[1], 1) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])})
    pose = ctx.pose()[:3, 3].copy()
    ctx.close()
print(os.environ.get("TFB_LIB_PATH", "default"), "frames/s:", " ".join("%.0f" % v for v in res), "| final t", pose)
