"""Oracle pinned to golden vectors.  The reference ships no tests or fixtures, so the
vectors under tests/golden/ were produced by tests/golden/make_golden.py from oracle/_ref — the build whose per-pixel /
per-voxel functions are the reference's own headers compiled from /root/reference — and the oracle port must
reproduce them bit for bit.  Runs on CPU."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import tfo

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("refgpu_"))   # refgpu_*: written by the reference's GPU build, used by the -m gpu tests


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _oracle_for(z, lib):
    mode, voxel, mu = z["params"]
    intr = z["intr"]
    rows, cols = z["depth"].shape[1:]
    return tfo.Oracle(lib=lib, cols=cols, rows=rows, fx=float(intr[0]), fy=float(intr[1]), cx=float(intr[2]), cy=float(intr[3]),
                      corrected_mode=int(mode), voxel_size=float(voxel), mu=float(mu))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_full_pipeline_matches_reference_vectors(path, oracle_lib):
    z = np.load(path)
    o = _oracle_for(z, oracle_lib)
    try:
        for i in range(len(z["depth"])):
            ok = o.process_frame(z["depth"][i])
            assert ok == bool(z["ok"][i]), i
            assert np.array_equal(o.pose().view(np.uint32), z["est_poses"][i].view(np.uint32)), f"pose of frame {i}"
            assert o.counters()["n_visible"] == z["n_visible"][i]
            assert o.voxel_updates() == z["voxel_updates"][i]
            if i == 1 and "f1_model_points_sha" in z:
                assert digest(o.level(3, 0)) == str(z["f1_model_points_sha"])
            elif i == 1:
                a, b = o.level(3, 0), z["f1_model_points"]
                assert np.array_equal(np.isnan(a), np.isnan(b))
                assert np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32))
                assert digest(o.raycast_result()) == str(z["f1_raycast_sha"])
        t = o.table()
        alloc = t[t["ptr"] >= 0]
        order = np.lexsort((alloc["pos"][:, 2], alloc["pos"][:, 1], alloc["pos"][:, 0]))
        alloc = alloc[order]
        assert np.array_equal(alloc["pos"], z["blocks"]), "allocated block set"
        vox = np.stack([o.block(int(p)) for p in alloc["ptr"]])
        assert digest(vox["sdf"]) == str(z["sdf_sha"]) and digest(vox["w"]) == str(z["w_sha"])
        assert digest(o.raycast_result()) == str(z["raycast_sha"])
        assert np.array_equal(np.sort(o.visible_ids()), z["vis_ids_sorted"])
    finally:
        o.close()


@pytest.mark.parametrize("path", [p for p in GOLDEN if "160x120_reference_mode" in p])
def test_stage_vectors(path, oracle_lib):
    z = np.load(path)
    L = oracle_lib
    d0, d1, intr = z["depth"][0], z["depth"][1], tuple(float(v) for v in z["intr"])
    assert digest(L.compute_dists(d0)) == str(z["st_dists_sha"])
    bf = L.bilateral(d0)
    assert np.array_equal(bf, z["st_bilateral"])
    tr = L.truncate_depth(bf, 2.0)
    assert np.array_equal(L.depth_pyr(tr), z["st_pyr1"])
    pts, nrm = L.points_normals(tr, intr)
    assert digest(pts) == str(z["st_points_sha"]) and digest(nrm) == str(z["st_normals_sha"])
    p1, n1 = L.points_normals(L.truncate_depth(L.bilateral(d1), 2.0), intr)
    v27, nc = L.icp_reduce(intr, z["st_icp_aff"], p1, n1, pts, nrm)
    assert nc == int(z["st_icp_ncorr"])
    assert np.array_equal(v27.view(np.uint32), z["st_icp27"].view(np.uint32))


@pytest.mark.skipif(not tfo.have_ref(), reason="oracle/_ref is only built where /root/reference exists")
def test_port_equals_reference_backed_build_live(s1_frames):
    """same inputs through both builds, every piece of state compared bit for bit (full 640x480)"""
    depth, _, _ = s1_frames
    res = {}
    for which in ("port", "ref"):
        o = tfo.Oracle(which=which)
        for i in range(4):
            assert o.process_frame(depth[i])
        res[which] = (o.table(), o.vis_type(), o.visible_ids(), o.pose(), o.raycast_result(), o.level(3, 0), o.level(4, 1), o.minmax())
        o.close()
    for a, b in zip(res["port"], res["ref"]):
        assert a.tobytes() == b.tobytes()


@pytest.mark.skipif(not tfo.have_ref(), reason="oracle/_ref is only built where /root/reference exists")
def test_mat4_inverse_port_equals_reference():
    P, R = tfo.Lib("port"), tfo.Lib("ref")
    rng = np.random.RandomState(3)
    for _ in range(200):
        m = rng.randn(16).astype(np.float32)
        ok1, a = P.mat4_inv_colmajor(m)
        ok2, b = R.mat4_inv_colmajor(m)
        assert ok1 == ok2 and a.tobytes() == b.tobytes()
    ok, _ = P.mat4_inv_colmajor(np.zeros(16, np.float32))
    assert not ok
