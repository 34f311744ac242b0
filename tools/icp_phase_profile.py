"""GPU box: phase timing of k_icp_all with a -DTFB_ICP_PROFILE build (clock64 stamps of CTA 0)"""
import sys, os, subprocess, ctypes as C
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from topfusion_b200 import capi, synth
depth, _, _ = synth.sequence("S1", 6)
ctx = capi.Context(corrected_mode=1)
for i in range(6):
    ctx.process_frame(depth[i])
out = np.zeros(512, np.int64)
rc = ctx.L.tfb_debug_icp_profile(out.ctypes.data_as(C.c_void_p))
t = out.reshape(64, 8)[:19, :6].astype(np.float64)
d = np.diff(t, axis=1) / 1.965e3  # us at 1965 MHz
print("iter  pixels  reduce  barrier  fold  solve   | total(us)")
for i in range(19):
    print("%3d  %6.2f  %6.2f  %6.2f  %6.2f  %6.2f   | %6.2f" % (i, *d[i], (t[i, 5] - t[i, 0]) / 1.965e3))
print("sum of iterations: %.1f us; span first..last %.1f us" % (d.sum(), (t[18, 5] - t[0, 0]) / 1.965e3))
cta = np.zeros(1024, np.int64)
if ctx.L.tfb_debug_icp_cta(cta.ctypes.data_as(C.c_void_p)) == 0:
    t = cta.reshape(256, 4)[:148].astype(np.float64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1000.0
    print("iteration 12 (level 0), all 148 CTAs, us since the first CTA started its pixels:")
    for name, col in (("pixel start", 0), ("pixel end", 1), ("row stored", 2), ("fold done", 3)):
        print("  %-12s min %5.2f  median %5.2f  p90 %5.2f  max %5.2f (CTA %d)" % (name, t[:, col].min(), np.median(t[:, col]), np.percentile(t[:, col], 90), t[:, col].max(), int(t[:, col].argmax())))
    d = t[:, 1] - t[:, 0]
    print("  pixel phase per CTA: min %.2f median %.2f max %.2f; slowest CTAs:" % (d.min(), np.median(d), d.max()), np.argsort(-d)[:8].tolist())
ctx.close()
