import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import tfo
from topfusion_b200 import capi, synth
for (cols, rows) in ((160, 120), (320, 240)):
    depth, poses, intr = synth.sequence("S1", 3, cols, rows)
    kw = dict(cols=cols, rows=rows, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], corrected_mode=1)
    o = tfo.Oracle(**kw); g = capi.Context(**kw)
    o.process_frame(depth[0]); o.process_frame(depth[1]); o.preprocess(depth[2])
    for lvl in range(3):
        for which in (1, 2, 3, 4):
            g.set_level(which, lvl, o.level(which, lvl))
    ok_o, a_o = o.estimate_transform()
    ok_g, a_g = g.estimate_transform()
    print(cols, rows, ok_o, ok_g, "oracle t", a_o[:3, 3], "gpu t", a_g[:3, 3], "diff %.2e" % np.abs(a_o - a_g).max())
    # single reductions per level at identity and at the oracle's answer
    for lvl in (2, 1, 0):
        li = tuple(np.float32(v) / np.float32(1 << lvl) for v in intr)
        for aff in (np.eye(4, dtype=np.float32), a_o):
            o27, n = o.L.icp_reduce(li, aff, o.level(1, lvl), o.level(2, lvl), o.level(3, lvl), o.level(4, lvl))
            g27 = g.icp_reduce(li, aff, o.level(1, lvl), o.level(2, lvl), o.level(3, lvl), o.level(4, lvl))
            print("   L%d ncorr %5d  max|o27| %.3e  max diff %.3e" % (lvl, n, np.abs(o27).max(), np.abs(o27 - g27).max()))
    g.close(); o.close()
