import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import tfo
from topfusion_b200 import capi, synth
cols, rows = 160, 120
depth, poses, intr = synth.sequence("S1", 4, cols, rows)
kw = dict(cols=cols, rows=rows, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], corrected_mode=1)
o = tfo.Oracle(**kw); g = capi.Context(**kw)
for i in range(4):
    oo, og = o.process_frame(depth[i]), g.process_frame(depth[i])
    print("frame", i, oo, og, "dt=%.2e" % np.abs(o.pose()[:3, 3] - g.pose()[:3, 3]).max(), o.counters()["n_visible"], g.counters()["n_visible"])
    for lvl in range(3):
        for which, nm in ((1, "vcurr"), (3, "vprev"), (4, "nprev")):
            a, b = o.level(which, lvl), g.level(which, lvl)
            na, nb = np.isnan(a[..., 0]), np.isnan(b[..., 0])
            m = ~na & ~nb
            print("   L%d %-5s nan-mismatch %5d  maxdiff %.3e  valid %d" % (lvl, nm, int((na != nb).sum()), np.abs(a[m] - b[m]).max() if m.any() else 0, int(m.sum())))
    ro, rg = o.raycast_result(), g.raycast_result()
    print("   raycast equal:", np.array_equal(ro.view(np.uint32), rg.view(np.uint32)), " minmax equal:", np.array_equal(np.ascontiguousarray(o.minmax()[:rows//8, :cols//8]).view(np.uint32), g.minmax().view(np.uint32)))
