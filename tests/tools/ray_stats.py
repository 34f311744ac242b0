"""TEST INFRASTRUCTURE (uses the oracle) — CPU (oracle, test infrastructure): distribution of march lengths per ray and per 8x4 warp patch on the S1 orbit."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tfo
from topfusion_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
depth, _, _ = synth.sequence("S1", n)
L = tfo.Lib("port")
o = tfo.Oracle(lib=L, corrected_mode=1)
buf = np.zeros((480, 640, 3), np.int32)
for i in range(n):
    if i == n - 1:
        L.lib.tfo_debug_ray_stats(buf.ctypes.data_as(C.c_void_p), C.c_int(640))
    o.process_frame(depth[i])
L.lib.tfo_debug_ray_stats(None, C.c_int(0))
st, mi, tr = buf[..., 0], buf[..., 1], buf[..., 2]
print("rays with >=1 step:", int((st > 0).sum()), "of", st.size)
for name, a in (("steps", st), ("missing-block steps", mi), ("in-loop trilinear", tr)):
    v = a[st > 0]
    print(f"{name:22s} mean {v.mean():6.2f}  p50 {np.percentile(v,50):5.0f}  p90 {np.percentile(v,90):5.0f}  p99 {np.percentile(v,99):5.0f}  max {v.max():5d}")
# per warp patch (8 wide x 4 high): the warp runs as long as its longest ray
pm = st.reshape(120, 4, 80, 8).max(axis=(1, 3))
print("per-warp max steps: mean %.1f p50 %.0f p90 %.0f p99 %.0f max %d; sum of warp-max %d vs sum of ray steps/32 %d" %
      (pm.mean(), np.percentile(pm, 50), np.percentile(pm, 90), np.percentile(pm, 99), pm.max(), pm.sum(), st.sum() // 32))
# per CTA (16x8)
cm = st.reshape(60, 8, 40, 16).max(axis=(1, 3))
print("per-CTA max steps: mean %.1f max %d" % (cm.mean(), cm.max()))
ys, xs = np.unravel_index(np.argsort(pm.ravel())[-5:], pm.shape)
print("longest warps at (patch y, x):", list(zip(ys.tolist(), xs.tolist())), pm[ys, xs])
# the warps the GPU profile found slowest (tools/ray_profile.py on the same frame): per-lane step / missing / trilinear counts
for w in [int(a) for a in sys.argv[2:]]:
    cta, wi = w // 4, w % 4
    x0, y0 = (cta % 40) * 16 + (wi & 1) * 8, (cta // 40) * 8 + (wi >> 1) * 4
    print(f"warp {w}: pixels x {x0}..{x0+7}, y {y0}..{y0+3}")
    for yy in range(y0, y0 + 4):
        print("   ", " ".join(f"{st[yy,xx]:2d}/{mi[yy,xx]:2d}/{tr[yy,xx]}" for xx in range(x0, x0 + 8)))
