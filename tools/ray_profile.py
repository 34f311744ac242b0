"""GPU box: per-warp lifetime of k_raycast with the instrumented build (make -C topfusion_b200/csrc prof):
TFB_LIB_PATH=topfusion_b200/libtfusion_b200_prof.so python tools/ray_profile.py"""
import ctypes as C, os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from topfusion_b200 import capi, synth
depth, _, _ = synth.sequence("S1", 8)
ctx = capi.Context(corrected_mode=1)
for i in range(8):
    ctx.flush_l2()
    ctx.process_frame(depth[i])
out = np.zeros(3 * 16384, np.int64)
assert ctx.L.tfb_debug_ray_profile(out.ctypes.data_as(C.c_void_p), C.c_int(out.size)) == 0
p = out.reshape(-1, 3)[:9600]
end_ns, cyc, sm = p[:, 0], p[:, 1], p[:, 2]
us = cyc / 1965.0
start_ns = end_ns - us * 1000
t0 = start_ns.min()
print("kernel span %.1f us (first warp start .. last warp end)" % ((end_ns.max() - t0) / 1000))
print("warp lifetime us: mean %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f" % (us.mean(), *np.percentile(us, [50, 90, 99]), us.max()))
print("warp start us:    p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile((start_ns - t0) / 1000, [50, 90, 99, 100])))
order = np.argsort(-us)[:10]
for w in order:
    cta = w // 4
    print("warp %5d cta (%2d,%2d) sm %3d start %6.1f us life %6.1f us" % (w, cta % 40, cta // 40, sm[w], (start_ns[w] - t0) / 1000, us[w]))
busy = np.zeros(148)
for s in range(148):
    m = sm == s
    if m.any():
        busy[s] = (end_ns[m].max() - start_ns[m].min()) / 1000
print("per-SM span us: mean %.1f min %.1f max %.1f" % (busy.mean(), busy.min(), busy.max()))
# timeline: resident warps every 5 us
for t in range(0, int((end_ns.max() - t0) / 1000) + 5, 5):
    tt = t0 + t * 1000
    print("t=%3d us resident warps %5d" % (t, int(((start_ns <= tt) & (end_ns > tt)).sum())))
ctx.close()
