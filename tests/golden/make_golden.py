"""Generates tests/golden/*.npz from oracle/_ref — the hybrid whose per-pixel / per-voxel functions are the
reference's OWN headers compiled where they lie under /root/reference (oracle/Makefile).  Runs only in the
authoring container (the reference tree is not on the GPU box); the fixtures it writes are committed.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tfo  # noqa: E402
from topfusion_b200 import synth  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scene_digest(o) -> dict:
    t = o.table()
    alloc = t[t["ptr"] >= 0]
    order = np.lexsort((alloc["pos"][:, 2], alloc["pos"][:, 1], alloc["pos"][:, 0]))
    alloc = alloc[order]
    vox = np.stack([o.block(int(p)) for p in alloc["ptr"]]) if len(alloc) else np.zeros((0, 512), tfo.VOXEL_DTYPE)
    return {"blocks": alloc["pos"].astype(np.int16), "sdf_sha": digest(vox["sdf"]), "w_sha": digest(vox["w"]),
            "sdf_sum": int(vox["sdf"].astype(np.int64).sum()), "w_sum": int(vox["w"].astype(np.int64).sum())}


def run(name, seq, cols, rows, n, lite=False, **kw):
    assert tfo.have_ref(), "build oracle/_ref first (make -C oracle)"
    L = tfo.Lib("ref")
    assert L.impl_name() == "reference"
    depth, poses, intr = synth.sequence(seq, n, cols, rows)
    o = tfo.Oracle(lib=L, cols=cols, rows=rows, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], **kw)
    out = {"depth": depth, "gt_poses": poses.astype(np.float32), "intr": np.array(intr, np.float32)}
    est, oks, nvis, vupd = [], [], [], []
    for i in range(n):
        oks.append(o.process_frame(depth[i]))
        est.append(o.pose().copy()); nvis.append(o.counters()["n_visible"]); vupd.append(o.voxel_updates())
        if i == 1:
            if lite:
                out["f1_model_points_sha"] = digest(o.level(3, 0))
            else:
                out["f1_model_points"] = o.level(3, 0)
            out["f1_raycast_sha"] = digest(o.raycast_result())
    out.update(est_poses=np.stack(est), ok=np.array(oks), n_visible=np.array(nvis), voxel_updates=np.array(vupd))
    sd = scene_digest(o)
    out.update(blocks=sd["blocks"], sdf_sha=sd["sdf_sha"], w_sha=sd["w_sha"], sdf_sum=sd["sdf_sum"], w_sum=sd["w_sum"])
    out["raycast_sha"] = digest(o.raycast_result())
    out["vis_ids_sorted"] = np.sort(o.visible_ids())
    out["params"] = np.array([kw.get("corrected_mode", 0), kw.get("voxel_size", 0.005), kw.get("mu", 0.02)], np.float64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if lite:   # full-resolution fixture: frames + per-frame results + digests only
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB; blocks", len(sd["blocks"]), "ok", oks, "nvis", nvis)
        o.close()
        return
    # stage-level vectors for frame 0/1 inputs
    d0 = depth[0]
    out["st_dists_sha"] = digest(L.compute_dists(d0))
    bf = L.bilateral(d0); out["st_bilateral"] = bf
    tr = L.truncate_depth(bf, 2.0); py = L.depth_pyr(tr); out["st_pyr1"] = py
    pts, nrm = L.points_normals(tr, intr); out["st_points_sha"] = digest(pts); out["st_normals_sha"] = digest(nrm)
    p1, n1 = L.points_normals(L.truncate_depth(L.bilateral(depth[1]), 2.0), intr)
    aff = L.rodrigues([0.001, -0.002, 0.0015], [0.003, -0.001, 0.002])
    v27, nc = L.icp_reduce(intr, aff, p1, n1, pts, nrm)
    out["st_icp_aff"] = aff; out["st_icp27"] = v27; out["st_icp_ncorr"] = nc
    np.savez_compressed(path, **out)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB; blocks", len(sd["blocks"]), "ok", oks, "nvis", nvis)
    o.close()


if __name__ == "__main__":
    run("s1_160x120_reference_mode", "S1", 160, 120, 5)
    run("s1_160x120_corrected_mode", "S1", 160, 120, 6, corrected_mode=1)
    run("s0_160x120_reference_8mm", "S0", 160, 120, 4, voxel_size=0.008)
    # full resolution: the only size at which frame-to-frame ICP is stiff enough for a 1e-4 GPU-vs-golden pose bound
    run("s1_640x480_corrected_mode", "S1", 640, 480, 3, lite=True, corrected_mode=1)
