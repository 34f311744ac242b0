"""TEST INFRASTRUCTURE (uses the oracle) — diagnostic (GPU box): per-frame GPU-vs-oracle differences of the full pipeline"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import tfo
from topfusion_b200 import capi, synth

def rot(Ra, Rb):
    D = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))

for seq, mode, n in (("S0", 0, 14), ("S1", 0, 8), ("S1", 1, 14)):
    depth, poses, intr = synth.sequence(seq, n)
    o = tfo.Oracle(corrected_mode=mode); g = capi.Context(corrected_mode=mode)
    print(f"== {seq} corrected={mode}")
    for i in range(n):
        oo, og = o.process_frame(depth[i]), g.process_frame(depth[i])
        po, pg = o.pose(), g.pose()
        # compare maps
        d0o, d0g = o.level(0, 0), g.level(0, 0)
        vpo, vpg = o.level(3, 0), g.level(3, 0)
        m = ~np.isnan(vpo[..., 0]) & ~np.isnan(vpg[..., 0])
        mism = int((np.isnan(vpo[..., 0]) != np.isnan(vpg[..., 0])).sum())
        print(i, oo, og, "dt=%.2e dr=%.2e" % (np.abs(po[:3, 3] - pg[:3, 3]).max(), rot(po[:3, :3], pg[:3, :3])),
              "depth0 diff px=%d" % int((d0o != d0g).sum()), "model-map max diff=%.2e nan-mismatch=%d" % (np.abs(vpo[m] - vpg[m]).max() if m.any() else 0, mism),
              "nvis", o.counters()["n_visible"], g.counters()["n_visible"], "t=", np.round(pg[:3, 3], 4), "gt=", np.round(poses[i][:3, 3], 4))
    g.close(); o.close()
