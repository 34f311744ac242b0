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
g = out.reshape(64, 8)[60, :4].astype(np.float64)
if g[0] > 0:
    print("CTA 0, globaltimer: kernel entry -> first iteration %.2f us; first -> end of last iteration %.2f us; epilogue (type 3, pose, host publish) %.2f us"
          % ((g[1] - g[0]) / 1e3, (g[2] - g[1]) / 1e3, (g[3] - g[2]) / 1e3))
e = out.reshape(64, 8)[61, :4].astype(np.float64)
if e[0] > 0:
    print("epilogue of CTA 0 (clock64): result + pose matrices %.2f us, state block -> pinned host %.2f us, system fence + sequence word %.2f us"
          % ((e[1] - e[0]) / 1.965e3, (e[2] - e[1]) / 1.965e3, (e[3] - e[2]) / 1.965e3))
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
pix = np.zeros(256 * 24, np.int32)
if hasattr(ctx.L, "tfb_debug_icp_pix") and ctx.L.tfb_debug_icp_pix(pix.ctypes.data_as(C.c_void_p)) == 0:
    p = pix.reshape(256, 24)[:148]
    sm, nl, cyc = p[:, 0], p[:, 1], p[:, 2:21].astype(np.float64)
    us = cyc / 1965.0
    print("pixel phase per CTA (clock64 of thread 0), level-0 iterations 9..18: median %.2f us, p90 %.2f, max %.2f" % (np.median(us[:, 9:]), np.percentile(us[:, 9:], 90), us[:, 9:].max()))
    print("list length at level 0: min %d median %d max %d" % (nl.min(), np.median(nl), nl.max()))
    slow = np.argsort(-us[:, 9:].mean(axis=1))[:10]
    for c in slow:
        print("  CTA %3d SM %3d list %4d : " % (c, sm[c], nl[c]) + " ".join("%.2f" % v for v in us[c, 9:]))
    # is slowness a property of the CTA (persistent) or of the iteration (random)?
    r = np.corrcoef(us[:, 10], us[:, 15])[0, 1]
    print("correlation of a CTA's pixel time between iterations 10 and 15: %.2f; with its list length: %.2f" % (r, np.corrcoef(us[:, 9:].mean(axis=1), nl)[0, 1]))
    print("per-iteration max over CTAs: " + " ".join("%.2f" % v for v in us.max(axis=0)))
    print("per-iteration median       : " + " ".join("%.2f" % v for v in np.median(us, axis=0)))
ctx.close()
