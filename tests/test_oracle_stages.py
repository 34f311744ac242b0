"""Host-side checks of the oracle's stages: edge cases the domain has (empty / ragged frames, cut-offs, image
borders) and properties that pin the restated semantics of src/cuda/imgproc.cu.  Runs on CPU."""
import numpy as np

from topfusion_b200 import synth


def test_synthetic_sequences_are_deterministic():
    a, pa, ia = synth.sequence("S1", 2)
    b, pb, ib = synth.sequence("S1", 2)
    assert np.array_equal(a, b) and np.array_equal(pa, pb) and ia == ib
    assert a.dtype == np.uint16 and a.shape == (2, 480, 640)
    assert 0.4 < (a[0] > 0).mean() < 0.7 and a.max() <= 10000
    assert np.allclose(pa[0], np.eye(4))
    h, ph, _ = synth.sequence("S0", 30)
    assert np.abs(ph[:, :3, 3]).max() < 0.005


def test_compute_dists_cutoffs(oracle_lib):
    d = np.array([[0, 1, 2046, 2047, 2048, 65535]], np.uint16)
    out = oracle_lib.compute_dists(d)
    assert np.array_equal(out, np.array([[-1, 0.001, 2.046, -1, -1, -1]], np.float32))   # imgproc.cu:277


def test_bilateral_window_is_edge_exclusive(oracle_lib):
    """window [x-3, min(x+4, cols-1)): the last column/row never contributes (imgproc.cu:26-27)"""
    d = np.full((32, 32), 1000, np.uint16)
    d[:, -1] = 1040
    d[-1, :] = 1040
    out = oracle_lib.bilateral(d)
    assert np.all(out[:-1, :-1] == 1000)
    z = np.zeros((16, 24), np.uint16)
    assert not oracle_lib.bilateral(z).any()


def test_truncate_and_pyramid(oracle_lib):
    d = np.array([[1999, 2000, 2001, 0]], np.uint16)
    assert np.array_equal(oracle_lib.truncate_depth(d, 2.0), [[1999, 2000, 0, 0]])
    src = np.full((16, 16), 1000, np.uint16)
    src[4:8, 4:8] = 1300           # more than 3*sigma (120 mm) away from its surroundings: never mixed in
    p = oracle_lib.depth_pyr(src)
    assert p.shape == (8, 8) and set(np.unique(p)) <= {1000, 1300}
    src2 = np.full((16, 16), 1000, np.uint16); src2[::2, ::2] = 1010
    assert np.all((oracle_lib.depth_pyr(src2) >= 1000) & (oracle_lib.depth_pyr(src2) <= 1010))
    assert not oracle_lib.depth_pyr(np.zeros((8, 8), np.uint16)).any()


def test_points_normals_invalid_rules(oracle_lib):
    intr = synth.DEFAULT_INTR
    d = np.full((8, 8), 1000, np.uint16)
    d[3, 3] = 0
    pts, nrm = oracle_lib.points_normals(d, intr)
    assert np.isnan(pts[-1]).all() and np.isnan(pts[:, -1]).all()          # last row / column
    assert np.isnan(pts[3, 3]).all() and np.isnan(pts[3, 2]).all() and np.isnan(pts[2, 3]).all()   # any of the 3 depths is 0
    ok = ~np.isnan(pts[..., 0])
    assert np.allclose(pts[ok][:, 2], 1.0) and np.all(pts[ok][:, 3] == 1.0)
    assert np.allclose(nrm[ok][:, :3], [0, 0, -1], atol=1e-6)            # fronto-parallel plane faces the camera
    p2, n2 = oracle_lib.resize_points_normals(pts, nrm)
    assert p2.shape == (4, 4, 4)
    assert np.isnan(p2[1, 1, 0]) and p2[1, 1, 3] == 0.0                  # NaN propagates, w = 0
    assert p2[0, 0, 3] == 1.0 and n2[0, 0, 3] == 0.0


def test_empty_frame_tracking_fails_and_resets():
    from oracle import tfo
    depth, _, _ = synth.sequence("S1", 2, 160, 120)
    intr = synth.intrinsics_for(160, 120)
    o = tfo.Oracle(cols=160, rows=120, fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3])
    try:
        assert o.process_frame(depth[0])
        assert not o.process_frame(np.zeros_like(depth[0]))       # no correspondences -> det = 0 -> reset
        c = o.counters()
        assert c["resets"] == 1 and c["frame_counter"] == 0 and c["n_allocated"] == 0
        assert o.num_poses() == 1 and np.array_equal(o.pose(), np.eye(4, dtype=np.float32))
        assert o.process_frame(depth[1])
    finally:
        o.close()
