"""Generates tests/golden/refgpu_*.npz by running the REFERENCE'S OWN GPU CODE on a B200.

The reference library (/root/reference/tfusion) is device-only for imgproc.cu and proj_icp.cu, so nothing in the
authoring container can execute those kernels.  baseline/ref_gpu/ builds the reference's sources — patched only as far as
patch_ref.py lists so that they compile with CUDA 12.9 for sm_100a, with the reference's own nvcc flags — into
baseline/_ref/libref_gpu_nodebug.so; that library travels to the GPU box with the snapshot, and this script calls it there:

    gpurun -- 'python tests/golden/make_refgpu_golden.py'      # writes tests/golden/ AND gpurun_out/refgpu_golden/

Every array below is the output of a reference kernel / engine method; inputs are stored beside the outputs so the tests
never depend on regenerating them.  Fixtures pin SURVEY §8 rows a1-a5, a7, a8, a17 (device-only arithmetic: __expf, rsqrt,
__fdividef, --prec-div=false, fused multiply-adds) and add device-side evidence for a6, a9-a16.
"""
import hashlib
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baseline.ref_gpu import refgpu  # noqa: E402
from topfusion_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MIRROR = os.path.join(ROOT, "gpurun_out", "refgpu_golden")


def sha(a):
    """digest with every NaN payload canonicalised (the reference writes quiet NaNs; only NaN-ness is specified)"""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = np.where(np.isnan(a), np.float32(np.nan), a).astype(np.float32)
    return hashlib.sha256(a.tobytes()).hexdigest()


def ragged(depth):
    """holes, an invalid band and out-of-range values: exercises every cut-off of imgproc.cu"""
    d = depth.copy()
    d[100:140, 200:260] = 0
    d[300:310, :] = 2047
    d[10:20, 10:50] = 9000
    return d


def rodrigues(rvec, t):
    L = refgpu.lib()
    out = np.zeros((4, 4), np.float32)
    L.refgpu_cv_affine(refgpu._p(np.asarray(rvec, np.float32)), refgpu._p(np.asarray(t, np.float32)), refgpu._p(out))
    return out


def pose_inv(m):
    L = refgpu.lib()
    out = np.zeros((4, 4), np.float32)
    L.refgpu_cv_affine_inv(refgpu._p(np.ascontiguousarray(m, np.float32)), refgpu._p(out))
    return out


def save(name, **arrs):
    os.makedirs(MIRROR, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    shutil.copy(path, os.path.join(MIRROR, name + ".npz"))
    print(f"{name}: {os.path.getsize(path) // 1024} KiB, {len(arrs)} arrays")


def stages():
    """a1-a5, a17, a7/a8, a6 at 640x480"""
    depth, gt, intr = synth.sequence("S1", 8)
    out = {"intr": np.array(intr, np.float32), "describe": np.array(refgpu.lib().refgpu_describe().decode())}
    d_in = ragged(depth[0])
    r = refgpu.stage_imgproc(d_in, intr)
    out.update(depth_in=d_in, dists_sha=np.array(sha(r["dists"])), dists_sample=r["dists"][::8, ::8].copy(), bilateral=r["bilateral"],
               depth0=r["depth"][0], depth1=r["depth"][1], depth2=r["depth"][2])
    for l in range(3):
        out[f"points{l}_sha"] = np.array(sha(r["points"][l]))
        out[f"normals{l}_sha"] = np.array(sha(r["normals"][l]))
    # full maps of the coarsest level, strided samples of the others (float32: the comparison is at 1e-5 and below)
    out.update(points2=r["points"][2], normals2=r["normals"][2], normals1_s3=r["normals"][1][::3, ::3].copy(), points1_s3=r["points"][1][::3, ::3].copy(),
               normals0_s5=r["normals"][0][::5, ::5].copy(), points0_s5=r["points"][0][::5, ::5].copy())
    # a17: resize of the level-2 maps (any size works; level 2 keeps the stored input small)
    rp, rn = refgpu.stage_resize(r["points"][2], r["normals"][2])
    out.update(resize_points=rp, resize_normals=rn)
    rp1, rn1 = refgpu.stage_resize(r["points"][0], r["normals"][0])
    out.update(resize_l0_points_sha=np.array(sha(rp1)), resize_l0_normals_s3=rn1[::3, ::3].copy())

    # a7/a8: the 27-vector at fixed transforms.  Maps of frames 0 (model) and 3 (current), the reference's own.
    ma = refgpu.stage_imgproc(depth[0], intr)
    mb = refgpu.stage_imgproc(depth[3], intr)
    out.update(icp_depth_model=depth[0], icp_depth_curr=depth[3],
               icp_model_points2=ma["points"][2], icp_model_normals2=ma["normals"][2], icp_curr_points2=mb["points"][2], icp_curr_normals2=mb["normals"][2])
    affs = {"identity": np.eye(4, dtype=np.float32), "small": rodrigues([0.002, -0.003, 0.001], [0.004, -0.002, 0.003]),
            "orbit3": rodrigues([0.0, 0.026, 0.0], [0.03, 0.0, 0.001]), "far": rodrigues([0, 0, 0], [0, 0, 5.0])}
    for name, aff in affs.items():
        out[f"icp_aff_{name}"] = aff
        for l in range(3):
            out[f"icp27_L{l}_{name}"] = refgpu.stage_icp_sums(intr, aff, mb["points"][l], mb["normals"][l], ma["points"][l], ma["normals"][l], level=l)
    # a6 + a7/a8 as a loop: ProjectiveICP::estimateTransform frame 0 -> k
    for k in (1, 3, 6):
        mk = refgpu.stage_imgproc(depth[k], intr)
        ok, aff = refgpu.stage_estimate(intr, mk["points"], mk["normals"], ma["points"], ma["normals"])
        out[f"est_depth_{k}"] = depth[k]
        out[f"est_ok_{k}"] = np.array(ok)
        out[f"est_affine_{k}"] = aff
    save("refgpu_stages_640x480", **out)


def block_positions(table):
    alloc = table[table["ptr"] >= 0]
    order = np.lexsort((alloc["pos"][:, 2], alloc["pos"][:, 1], alloc["pos"][:, 0]))
    return alloc[order]


def scene():
    """a9-a16 with injected poses: allocate + integrate twice at pose A (the second pass admits the same-hash losers of the
    reference's racy first pass), once at pose B; then expected depths, raycast, model maps at pose B."""
    depth, gt, intr = synth.sequence("S1", 14)
    ia, ib = 10, 12          # generic (non axis-aligned) poses: knife-edge voxels are rare
    pa, pb = gt[ia].astype(np.float32), gt[ib].astype(np.float32)
    R = refgpu.RefTopFu()
    out = {"intr": np.array(intr, np.float32), "depth_a": depth[ia], "depth_b": depth[ib], "pose_a": pa, "pose_b": pb,
           "pose_a_w2c": pose_inv(pa), "pose_b_w2c": pose_inv(pb)}
    R.scene_integrate(depth[ia], pa)
    t1 = block_positions(R.table())
    out["blocks_pass1"] = t1["pos"].copy()
    out["counters_pass1"] = np.array(list(R.counters().values()), np.int32)
    R.scene_integrate(depth[ia], pa)
    t2 = block_positions(R.table())
    out["blocks_pass2"] = t2["pos"].copy()
    R.scene_integrate(depth[ib], pb)
    t3 = block_positions(R.table())
    vox = R.voxels()
    out["blocks"] = t3["pos"].copy()
    out["sdf"] = np.stack([vox[p]["sdf"] for p in t3["ptr"]])
    out["w"] = np.stack([vox[p]["w"] for p in t3["ptr"]])
    ids, types = R.visible()
    tab = R.table()
    vis = tab[ids]
    vis = vis[vis["ptr"] >= 0]["pos"]
    out["visible_blocks"] = vis[np.lexsort((vis[:, 2], vis[:, 1], vis[:, 0]))].copy()
    out["counters"] = np.array(list(R.counters().values()), np.int32)
    R.scene_raycast(pb)
    rc = R.raycast_result()
    rng = R.range_image()
    out["range_image"] = rng[: 480 // 8, : 640 // 8].copy()      # only this corner is written
    out["raycast_sha"] = np.array(sha(rc))
    out["raycast_s4"] = rc[::4, ::4].copy()
    out["raycast_hit_mask"] = np.packbits(rc[..., 3] > 0)
    for l in range(3):
        p, n = R.maps(1, l)
        out[f"model_points{l}_sha"] = np.array(sha(p)); out[f"model_normals{l}_sha"] = np.array(sha(n))
    p2, n2 = R.maps(1, 2)
    p0, n0 = R.maps(1, 0)
    out.update(model_points2=p2, model_normals2=n2, model_points0_s5=p0[::5, ::5].copy(), model_normals0_s5=n0[::5, ::5].copy())
    ids2, _ = R.visible()
    tab2 = R.table()
    vis2 = tab2[ids2]; vis2 = vis2[vis2["ptr"] >= 0]["pos"]
    out["visible_after_raycast_n"] = np.array(len(ids2))
    out["render_at_b_sha"] = np.array(sha(R.render_at(pb)))
    out["render_at_b_s4"] = R.render_at(pb)[::4, ::4].copy()
    R.close()

    # identity pose (frame 0 of every run): voxel planes sit exactly on the truncation boundary eta == -mu for every depth that
    # is a multiple of 5 mm; which side the device lands on depends on its fused multiply-adds.  Statistics only.
    R = refgpu.RefTopFu()
    eye = np.eye(4, dtype=np.float32)
    R.scene_integrate(depth[0], eye)
    R.scene_integrate(depth[0], eye)
    t = block_positions(R.table())
    vox = R.voxels()
    out["identity_blocks"] = t["pos"].copy()
    out["identity_sdf"] = np.stack([vox[p]["sdf"] for p in t["ptr"]])
    out["identity_w"] = np.stack([vox[p]["w"] for p in t["ptr"]])
    out["identity_depth"] = depth[0]
    R.close()
    save("refgpu_scene_640x480", **out)


def pipeline():
    """a19: TopFu::operator() over the first frames of S1 and S0 — poses, verdicts, block counts per frame"""
    out = {}
    for seq, n in (("S1", 12), ("S0", 14)):
        dep, _, intr = synth.sequence(seq, n)
        R = refgpu.RefTopFu()
        poses, oks, nblocks, nvis = [], [], [], []
        for i in range(n):
            oks.append(R.frame(dep[i])); poses.append(R.pose())
            t = R.table(); nblocks.append(int((t["ptr"] >= 0).sum())); nvis.append(R.counters()["n_visible"])
        out[f"{seq}_depth_sha"] = np.array(sha(dep))
        out[f"{seq}_poses"] = np.stack(poses); out[f"{seq}_ok"] = np.array(oks); out[f"{seq}_nblocks"] = np.array(nblocks)
        out[f"{seq}_nvis"] = np.array(nvis)
        R.close()
        # run-to-run determinism of the reference itself (allocation races)
        R2 = refgpu.RefTopFu()
        p2 = []
        for i in range(n):
            R2.frame(dep[i]); p2.append(R2.pose())
        out[f"{seq}_poses_rerun"] = np.stack(p2)
        R2.close()
    save("refgpu_pipeline_640x480", **out)


if __name__ == "__main__":
    if not refgpu.available():
        raise SystemExit("baseline/_ref/libref_gpu_nodebug.so is missing: run `make -C baseline/ref_gpu` where /root/reference exists")
    stages()
    scene()
    pipeline()
