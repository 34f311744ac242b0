"""Runs on the B200: the reference's own (patched-to-compile) GPU path beside ours, stage by stage and as a whole
pipeline.  Prints difference statistics and writes the raw reference outputs to gpurun_out/ for study.

    python tools/refgpu_explore.py [--frames 40] [--time-frames 100]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline.ref_gpu import refgpu  # noqa: E402
from topfusion_b200 import capi, synth  # noqa: E402
from oracle import tfo  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def ragged(depth):
    d = depth.copy()
    d[100:140, 200:260] = 0
    d[300:310, :] = 2047
    d[10:20, 10:50] = 9000
    return d


def fstat(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    nan_mismatch = int((na != nb).sum())
    m = ~(na | nb)
    bits_equal = int((a[m].view(np.uint32) == b[m].view(np.uint32)).sum()) if a.dtype == np.float32 else int((a[m] == b[m]).sum())
    d = np.abs(a[m].astype(np.float64) - b[m].astype(np.float64))
    rec = {"n": int(m.sum()), "nan_mismatch": nan_mismatch, "bit_equal_frac": bits_equal / max(1, int(m.sum())),
           "max_abs": float(d.max()) if d.size else 0.0, "mean_abs": float(d.mean()) if d.size else 0.0}
    print(f"  {name:34s} n={rec['n']:8d} nan_mismatch={nan_mismatch:6d} bit_equal={rec['bit_equal_frac']:.6f} max_abs={rec['max_abs']:.3e} mean_abs={rec['mean_abs']:.3e}")
    return rec


def rot_angle(Ra, Rb):
    D = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--time-frames", type=int, default=100)
    args = ap.parse_args()
    report = {"reference": refgpu.lib("nodebug").refgpu_describe().decode()}
    print(report["reference"])
    tfo.build()
    OL = tfo.Lib("port")
    depth, gt, intr = synth.sequence("S1", max(args.frames, 8))
    d0 = ragged(depth[0])
    ctx = capi.Context()

    # ---- stage level -------------------------------------------------------------------------------
    print("== imgproc stages: reference GPU vs ours (G) and vs oracle port (O)")
    ref = refgpu.stage_imgproc(d0, intr)
    st = {}
    g_d = ctx.compute_dists(d0); o_d = OL.compute_dists(d0)
    st["dists_G"] = fstat("dists ref-vs-G", ref["dists"], g_d); st["dists_O"] = fstat("dists ref-vs-O", ref["dists"], o_d)
    g_b = ctx.bilateral(d0); o_b = OL.bilateral(d0)
    st["bilateral_G"] = fstat("bilateral ref-vs-G", ref["bilateral"].astype(np.int32), g_b.astype(np.int32))
    st["bilateral_O"] = fstat("bilateral ref-vs-O", ref["bilateral"].astype(np.int32), o_b.astype(np.int32))
    # downstream stages are fed the REFERENCE's own upstream output so that each stage is judged alone
    rb = ref["bilateral"]
    g_t = ctx.truncate_depth(rb, 2.0)
    st["trunc_G"] = fstat("truncate ref-vs-G", ref["depth"][0].astype(np.int32), g_t.astype(np.int32))
    for l in (1, 2):
        g_p = ctx.depth_pyr(ref["depth"][l - 1]); o_p = OL.depth_pyr(ref["depth"][l - 1])
        st[f"pyr{l}_G"] = fstat(f"pyramid L{l} ref-vs-G", ref["depth"][l].astype(np.int32), g_p.astype(np.int32))
        st[f"pyr{l}_O"] = fstat(f"pyramid L{l} ref-vs-O", ref["depth"][l].astype(np.int32), o_p.astype(np.int32))
    for l in range(3):
        li = tuple(np.float32(v) / np.float32(1 << l) for v in intr)
        gp, gn = ctx.points_normals(ref["depth"][l], li)
        op, on = OL.points_normals(ref["depth"][l], li)
        st[f"points{l}_G"] = fstat(f"points L{l} ref-vs-G", ref["points"][l], gp)
        st[f"normals{l}_G"] = fstat(f"normals L{l} ref-vs-G", ref["normals"][l], gn)
        st[f"points{l}_O"] = fstat(f"points L{l} ref-vs-O", ref["points"][l], op)
        st[f"normals{l}_O"] = fstat(f"normals L{l} ref-vs-O", ref["normals"][l], on)
    rp, rn = refgpu.stage_resize(ref["points"][0], ref["normals"][0])
    gp, gn = ctx.resize_points_normals(ref["points"][0], ref["normals"][0])
    st["resize_points_G"] = fstat("resize points ref-vs-G", rp, gp)
    st["resize_normals_G"] = fstat("resize normals ref-vs-G", rn, gn)
    report["stages"] = st

    # ---- ICP sums at fixed transforms -----------------------------------------------------------------
    print("== ICP 27-vector at fixed transforms (maps: the reference's own, frames 0 and 3)")
    ra = refgpu.stage_imgproc(depth[0], intr)
    rb_ = refgpu.stage_imgproc(depth[3], intr)
    icp = {}
    affs = {"identity": np.eye(4, dtype=np.float32), "small": OL.rodrigues([0.002, -0.003, 0.001], [0.004, -0.002, 0.003]),
            "orbit3": OL.rodrigues([0.0, 0.026, 0.0], [0.03, 0.0, 0.001])}
    sums_keep = {}
    for l in range(3):
        li = tuple(np.float32(v) / np.float32(1 << l) for v in intr)
        for name, aff in affs.items():
            r27 = refgpu.stage_icp_sums(intr, aff, rb_["points"][l], rb_["normals"][l], ra["points"][l], ra["normals"][l], level=l)
            g27 = ctx.icp_reduce(li, aff, rb_["points"][l], rb_["normals"][l], ra["points"][l], ra["normals"][l])
            o27, nc = OL.icp_reduce(li, aff, rb_["points"][l], rb_["normals"][l], ra["points"][l], ra["normals"][l])
            sc = float(np.abs(r27).max()) or 1.0
            icp[f"L{l}_{name}"] = {"scale": sc, "G_rel": float(np.abs(g27 - r27).max() / sc), "O_rel": float(np.abs(o27 - r27).max() / sc),
                                   "O_bit_equal": int((o27.view(np.uint32) == r27.view(np.uint32)).sum()), "ncorr_oracle": int(nc)}
            sums_keep[f"icp27_L{l}_{name}"] = r27
            print(f"  L{l} {name:9s} scale={sc:.4e} G_rel={icp[f'L{l}_{name}']['G_rel']:.3e} O_rel={icp[f'L{l}_{name}']['O_rel']:.3e} "
                  f"O_bit_equal={icp[f'L{l}_{name}']['O_bit_equal']}/27 ncorr={nc}")
    report["icp_sums"] = icp

    # ---- whole estimateTransform ----------------------------------------------------------------------
    print("== estimateTransform (frames 0 -> k), reference GPU vs ours")
    est = {}
    for k in (1, 3, 6):
        rk = refgpu.stage_imgproc(depth[k], intr)
        ok_r, a_r = refgpu.stage_estimate(intr, rk["points"], rk["normals"], ra["points"], ra["normals"])
        g = capi.Context()
        g.preprocess(depth[0])
        for lvl in range(3):
            g.set_level(3, lvl, ra["points"][lvl]); g.set_level(4, lvl, ra["normals"][lvl])
        g.preprocess(depth[k])
        for lvl in range(3):   # current maps: the reference's, so only the ICP differs
            g.set_level(1, lvl, rk["points"][lvl]); g.set_level(2, lvl, rk["normals"][lvl])
        ok_g, a_g = g.estimate_transform()
        g.close()
        dt = float(np.abs(a_r[:3, 3] - a_g[:3, 3]).max()); dr = rot_angle(a_r[:3, :3], a_g[:3, :3])
        est[f"0->{k}"] = {"ok_ref": ok_r, "ok_ours": ok_g, "dt_m": dt, "dr_rad": dr, "t_ref": a_r[:3, 3].tolist()}
        print(f"  0->{k}: ok {ok_r}/{ok_g} |dt|={dt:.3e} m  dR={dr:.3e} rad  t_ref={a_r[:3,3]}")
    report["estimate"] = est

    # ---- whole pipeline, reference mode ---------------------------------------------------------------
    print("== TopFu::operator() sequence: reference GPU vs ours (reference mode) — S1 and S0")
    pipe = {}
    for seq in ("S1", "S0"):
        dep, _, _ = synth.sequence(seq, args.frames)
        R = refgpu.RefTopFu()
        G = capi.Context(corrected_mode=0, defer_tail=0)
        rows = []
        for i in range(args.frames):
            okr = R.frame(dep[i]); okg = G.process_frame(dep[i])
            pr, pg = R.pose(), G.pose()
            dt = float(np.abs(pr[:3, 3] - pg[:3, 3]).max()); dr = rot_angle(pr[:3, :3], pg[:3, :3])
            tr = R.table(); tg = G.table()
            sr = set(map(tuple, tr["pos"][tr["ptr"] >= 0].tolist())); sg = set(map(tuple, tg["pos"][tg["ptr"] >= 0].tolist()))
            rows.append({"i": i, "ok_ref": okr, "ok_ours": okg, "dt": dt, "dr": dr, "blocks_ref": len(sr), "blocks_ours": len(sg),
                         "symdiff": len(sr ^ sg), "nvis_ref": R.counters()["n_visible"], "nvis_ours": G.counters()["n_visible"],
                         "t_ref_norm": float(np.linalg.norm(pr[:3, 3]))})
            print(f"  {seq} f{i:02d} ok {int(okr)}/{int(okg)} |dt|={dt:.2e} dR={dr:.2e} blocks {len(sr)}/{len(sg)} symdiff {len(sr ^ sg)} "
                  f"nvis {rows[-1]['nvis_ref']}/{rows[-1]['nvis_ours']} |t_ref|={rows[-1]['t_ref_norm']:.4f}")
            if i == 1:
                # voxel-level comparison after two integrations
                vr = R.voxels(); mr = {tuple(p): q for p, q in zip(tr["pos"][tr["ptr"] >= 0].tolist(), tr["ptr"][tr["ptr"] >= 0].tolist())}
                bad = 0; tot = 0; maxd = 0
                gb = G.blocks_by_pos()
                for pos, blk in gb.items():
                    if pos in mr:
                        rblk = vr[mr[pos]]
                        dd = np.abs(rblk["sdf"].astype(np.int32) - blk["sdf"].astype(np.int32))
                        bad += int((dd > 0).sum()) + int((rblk["w"] != blk["w"]).sum()); tot += 512; maxd = max(maxd, int(dd.max()))
                pipe[f"{seq}_voxels_f1"] = {"voxels": tot, "differing": bad, "max_sdf_diff_lsb": maxd}
                print(f"    voxels after frame 1: {tot} compared, {bad} differ, max |sdf diff| = {maxd} LSB (of 32767)")
                rr, rg = R.raycast_result(), G.raycast_result()
                pipe[f"{seq}_raycast_f1"] = fstat("raycast result ref-vs-G", rr, rg)
                mp_r, mn_r = R.maps(1, 0); mp_g = G.level(3, 0); mn_g = G.level(4, 0)
                pipe[f"{seq}_model_points_f1"] = fstat("model points L0 ref-vs-G", mp_r, mp_g)
                pipe[f"{seq}_model_normals_f1"] = fstat("model normals L0 ref-vs-G", mn_r, mn_g)
        pipe[seq] = rows
        R.close(); G.close()
    report["pipeline"] = pipe

    # ---- timing: reference GPU as shipped / without its debug work, same frames ----------------------------
    print("== timing, S1 orbit, reference mode (frames resident on the host, upload inside the loop as demo.cpp does)")
    n = args.time_frames
    dep, _, _ = synth.sequence("S1", n)
    timing = {}
    for variant in ("nodebug", "asis"):
        for rep in range(2):
            R = refgpu.RefTopFu(variant=variant)
            oks = 0
            for i in range(5):
                R.frame(dep[i])
            R.sync(); t0 = time.perf_counter()
            for i in range(5, n):
                oks += R.frame(dep[i])
            R.sync(); t1 = time.perf_counter()
            fps = (n - 5) / (t1 - t0)
            timing[f"{variant}_run{rep}"] = {"fps": fps, "ok_frames": oks, "frames": n - 5}
            print(f"  reference GPU {variant:8s} run {rep}: {fps:8.1f} frames/s ({oks}/{n - 5} frames tracked)")
            R.close()
    G = capi.Context(corrected_mode=0)
    for i in range(5):
        G.process_frame(dep[i])
    G.sync(); t0 = time.perf_counter(); oks = 0
    for i in range(5, n):
        oks += G.process_frame(dep[i])
    G.sync(); t1 = time.perf_counter()
    timing["ours_reference_mode"] = {"fps": (n - 5) / (t1 - t0), "ok_frames": oks, "frames": n - 5}
    print(f"  ours (reference mode, host frames, wall clock incl. Python): {(n - 5) / (t1 - t0):8.1f} frames/s ({oks} tracked)")
    G.close()
    report["timing"] = timing

    with open(os.path.join(OUT, "refgpu_explore.json"), "w") as f:
        json.dump(report, f, indent=1, default=str)
    np.savez_compressed(os.path.join(OUT, "refgpu_explore_stage.npz"), depth=d0, bilateral=ref["bilateral"], pyr1=ref["depth"][1], pyr2=ref["depth"][2],
                        dists=ref["dists"], points1=ref["points"][1], normals1=ref["normals"][1], points2=ref["points"][2], normals2=ref["normals"][2],
                        depth0=ref["depth"][0], **sums_keep)
    print("wrote gpurun_out/refgpu_explore.json")


if __name__ == "__main__":
    main()
