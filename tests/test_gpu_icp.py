"""GPU parity — projective ICP (src/cuda/proj_icp.cu, src/projective_icp.cpp).  Floating point with a
different (but fixed) summation order: fp32 shuffle partials + fp64 final fold here, fp32 256-wide trees in the
reference.  Tolerances: 27-vector 2e-4 relative to the largest |term| of its kind; transform 1e-4 m / 1e-4 rad
(the north-star pose tolerance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rot_angle(Ra, Rb):
    """angle of Ra^T Rb from its skew part (arccos of the trace has a 5e-4 noise floor on fp32 rotations)"""
    D = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


@pytest.fixture(scope="module")
def maps(oracle_lib, s0_frames):
    depth, _, intr = s0_frames
    def prep(d):
        d0 = oracle_lib.truncate_depth(oracle_lib.bilateral(d), 2.0)
        return oracle_lib.points_normals(d0, intr)
    return prep(depth[0]), prep(depth[6]), intr


def test_reduction_matches_for_fixed_transforms(gpu, oracle_lib, maps):
    (vp, np_), (vc, nc), intr = maps
    c = gpu.Context()
    try:
        # At the identity every current point projects exactly onto its own pixel centre, so the point-sampled
        # texel floor(coo) flips with the last bit of __fdividef (device) vs IEEE division (oracle): a knife edge
        # of the reference's own design, hence the looser bound there.
        cases = ((np.eye(4, dtype=np.float32), 2e-3),
                 (oracle_lib.rodrigues([0.002, -0.003, 0.001], [0.004, -0.002, 0.003]), 5e-4))
        for aff, tol in cases:
            o27, ncorr = oracle_lib.icp_reduce(intr, aff, vc, nc, vp, np_)
            g27 = c.icp_reduce(intr, aff, vc, nc, vp, np_)
            assert ncorr > 100000
            scale = np.abs(o27).max()
            assert np.abs(g27 - o27).max() <= tol * scale, (np.abs(g27 - o27).max(), scale)
    finally:
        c.close()


def test_reduction_no_correspondences_is_zero(gpu, oracle_lib, maps):
    (vp, np_), (vc, nc), intr = maps
    c = gpu.Context()
    try:
        far = np.eye(4, dtype=np.float32); far[2, 3] = 5.0   # push everything 5 m away: nothing within 0.1 m
        g27 = c.icp_reduce(intr, far, vc, nc, vp, np_)
        o27, n = oracle_lib.icp_reduce(intr, far, vc, nc, vp, np_)
        assert n == 0 and not o27.any() and not g27.any()
        nan = np.full_like(vc, np.nan)
        assert not c.icp_reduce(intr, np.eye(4, dtype=np.float32), nan, nan, vp, np_).any()
    finally:
        c.close()


def test_estimate_transform_matches_oracle(gpu, s0_frames):
    from oracle import tfo
    depth, _, _ = s0_frames
    o = tfo.Oracle()
    g = gpu.Context()
    try:
        # model maps = frame 0, current = frame 6, installed on both sides through the same path
        o.preprocess(depth[0]); g.preprocess(depth[0])
        for lvl in range(3):
            for w_src, w_dst in ((1, 3), (2, 4)):
                o.set_level(w_dst, lvl, o.level(w_src, lvl))
                g.set_level(w_dst, lvl, o.level(w_src, lvl))
        o.preprocess(depth[6]); g.preprocess(depth[6])
        ok_o, a_o = o.estimate_transform()
        ok_g, a_g = g.estimate_transform()
        assert ok_o and ok_g
        assert np.abs(a_o[:3, 3] - a_g[:3, 3]).max() < 1e-4
        assert _rot_angle(a_o[:3, :3], a_g[:3, :3]) < 1e-4
        assert np.abs(a_o[:3, 3]).max() > 1e-4   # a real motion was estimated
    finally:
        g.close(); o.close()


def test_tracking_failure_is_reported(gpu, s0_frames):
    """no correspondences -> A = 0 -> |det| < 1e-15 -> estimateTransform returns false (projective_icp.cpp:197-203)"""
    depth, _, _ = s0_frames
    g = gpu.Context()
    try:
        g.preprocess(depth[0])           # current maps valid, model maps still all-zero/invalid
        nan = np.full((480, 640, 4), np.nan, np.float32)
        g.set_level(3, 0, nan); g.set_level(4, 0, nan)
        for lvl in (1, 2):
            n = np.full((480 >> lvl, 640 >> lvl, 4), np.nan, np.float32)
            g.set_level(3, lvl, n); g.set_level(4, lvl, n)
        ok, _ = g.estimate_transform()
        assert not ok
    finally:
        g.close()


def test_persistent_icp_is_reproducible_over_many_launches(gpu, s0_frames):
    """k_icp_all synchronises its 148 CTAs through epoch-stamped partial rows that every CTA polls (no counter, no fence):
    each 64-bit word carries {sum, epoch} and is written / read by one single-copy-atomic access (the round-1 layout
    relied on 32-byte sectors being written as a unit).  A reader that ever accepted a stale or torn word would fold a sum of
    another iteration: the 19-iteration result would then differ between launches.  300 launches, bit for bit."""
    depth, _, _ = s0_frames
    g = gpu.Context()
    try:
        g.preprocess(depth[0])
        for lvl in range(3):
            g.set_level(3, lvl, g.level(1, lvl)); g.set_level(4, lvl, g.level(2, lvl))
        g.preprocess(depth[6])
        ok0, a0 = g.estimate_transform()
        assert ok0
        for _ in range(300):
            ok, a = g.estimate_transform()
            assert ok and np.array_equal(a.view(np.uint32), a0.view(np.uint32))
    finally:
        g.close()


def test_icp_parameter_ranges_are_validated(gpu):
    """fixed-width encodings: 64 row epochs per launch -> at most 63 iterations per coarse-to-fine loop; 6 bits of
    step in the allocation claim key -> mu / voxel_size is bounded"""
    with pytest.raises(gpu.TfbError):
        gpu.Context(icp_iters=(40, 20, 10, 0))
    with pytest.raises(gpu.TfbError):
        gpu.Context(icp_iters=(10, -1, 4, 0))
    with pytest.raises(gpu.TfbError):
        gpu.Context(mu=0.5, voxel_size=0.002)
    c = gpu.Context(icp_iters=(30, 20, 13, 0))     # 63 in total: accepted
    c.close()


def test_frame_path_valid_pixel_list(gpu, s1_frames):
    """Level 0 of the frame path's ICP walks ONE ascending list of the pixels that hold a vertex (find_coresp skips a NaN vertex,
    proj_icp.cu:86-88), of which every CTA takes an equal share: validity bits from the level-0 map kernel, compacted by extra
    CTAs of the last preprocessing launch.  The list must be exactly the non-NaN vertices of the current level-0 map, in
    pixel order, frame after frame — and the poses of that path must be reproducible to the bit."""
    depth, _, _ = s1_frames
    a, b = gpu.Context(corrected_mode=1), gpu.Context(corrected_mode=1)
    try:
        assert a.icp_valid_list().size == 0          # nothing tracked yet
        for i in range(5):
            assert a.process_frame(depth[i]) and b.process_frame(depth[i])
            if i == 0:
                continue                              # frame 0 writes its maps straight into the model pyramid, no ICP
            want = np.flatnonzero(~np.isnan(a.level(1, 0)[..., 0].ravel())).astype(np.int32)
            got = a.icp_valid_list()
            assert got.size == want.size > 100000 and np.array_equal(got, want), i
            assert np.array_equal(a.pose().view(np.uint32), b.pose().view(np.uint32)), i
    finally:
        a.close(); b.close()
