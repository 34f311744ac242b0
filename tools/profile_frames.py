"""profiling driver (GPU box): N frames of the S1 orbit through tfb_process_frame_device, nothing else.
usage: python tools/profile_frames.py [n_frames] [mode] [cols rows voxel_mm]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from topfusion_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
depth, poses, intr = synth.sequence("S1", n)
ctx = capi.Context(corrected_mode=mode)
bufs = [ctx.upload(depth[i]) for i in range(n)]
for i in range(n):
    ok = ctx.process_frame_device(bufs[i])
print("done", ok, ctx.counters(), ctx.kernel_launches())
ctx.close()
