import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import tfo
from topfusion_b200 import capi, synth
cols, rows = 160, 120
depth, poses, intr = synth.sequence("S1", 3, cols, rows)
kw = dict(cols=cols, rows=rows, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], corrected_mode=1)
o = tfo.Oracle(**kw); g = capi.Context(**kw)
for i in range(2):
    o.process_frame(depth[i]); g.process_frame(depth[i])
o.preprocess(depth[2]); g.preprocess(depth[2])
names = {1: "vcurr", 2: "ncurr", 3: "vprev", 4: "nprev"}
for lvl in range(3):
    for which in (1, 2, 3, 4):
        a, b = o.level(which, lvl), g.level(which, lvl)
        same = ((a == b) | (np.isnan(a) & np.isnan(b)))
        bad = np.argwhere(~same.all(axis=-1))
        print("L%d %-5s differing pixels %4d" % (lvl, names[which], len(bad)), end="")
        if len(bad):
            y, x = bad[0]
            print("  e.g. (%d,%d) oracle %s gpu %s" % (y, x, a[y, x], b[y, x]), end="")
        print()
ok_o, a_o = o.estimate_transform(); ok_g, a_g = g.estimate_transform()
print("own maps: oracle t", a_o[:3, 3], "gpu t", a_g[:3, 3])
# feed the GPU's maps to the oracle
for lvl in range(3):
    for which in (1, 2, 3, 4):
        o.set_level(which, lvl, g.level(which, lvl))
ok_o2, a_o2 = o.estimate_transform()
print("oracle on GPU maps t", a_o2[:3, 3])
