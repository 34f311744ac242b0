"""The OpenCV calls the reference makes on the hot path are outside /root/reference (un-vendored dependency,
find_package(OpenCV 2.4.9 ...), CMakeLists.txt:18): cv::determinant / cv::solve(DECOMP_SVD) / cv::Affine3f(rvec,t) /
Affine3f product and inverse (projective_icp.cpp:197-209, topfu.cpp:243,281).  The oracle restates their published
algorithms; these tests pin the restatement against the cv2 4.13 build in this image.  Runs on CPU."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _spd6(rng, cond=1e3):
    q, _ = np.linalg.qr(rng.randn(6, 6))
    w = np.logspace(0, np.log10(cond), 6) * 50.0
    return ((q * w) @ q.T).astype(np.float32)


def test_solve_matches_cv2_svd(oracle_lib):
    rng = np.random.RandomState(0)
    for cond in (10, 1e3, 1e5):
        for _ in range(10):
            A = _spd6(rng, cond); A = ((A + A.T) / 2).astype(np.float32)
            b = rng.randn(6).astype(np.float32) * 10
            ok, x_cv = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)
            x = oracle_lib.solve6(A, b)
            ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
            scale = np.abs(ref).max()
            # the oracle solves in fp64: closer to the exact solution than OpenCV's fp32 Jacobi SVD
            assert np.abs(x - ref).max() <= 2e-6 * scale
            assert np.abs(x - x_cv.ravel()).max() <= max(5e-4 * cond / 1e3, 5e-5) * scale


def test_solve_rank_deficient(oracle_lib):
    """Singular A (a degenerate view): cv::solve(DECOMP_SVD) is documented to return the minimum-norm solution, but
    its fp32 Jacobi SVD leaves the null singular values at rounding-noise level, above its 2*FLT_EPSILON*sum(w)
    cut, so the answer carries an arbitrary null-space component (measured here: |x| ~ 80 instead of 0.76).  What IS
    well defined is pinned: both satisfy A x = b in the range space, and the oracle's answer is the true
    minimum-norm one.  Pose parity is therefore only claimed for full-rank systems."""
    rng = np.random.RandomState(1)
    v = rng.randn(6, 3)
    A = (v @ v.T).astype(np.float32) * 100          # rank 3
    b = (A.astype(np.float64) @ rng.randn(6)).astype(np.float32)
    _, x_cv = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)
    x = oracle_lib.solve6(A, b)
    x_np = np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=1e-6)[0]
    assert np.abs(x - x_np).max() < 1e-3 * max(1.0, np.abs(x_np).max())
    A64 = A.astype(np.float64)
    scale = np.abs(b).max()
    assert np.abs(A64 @ x.astype(np.float64) - b).max() < 1e-3 * scale
    assert np.abs(A64 @ x_cv.ravel().astype(np.float64) - b).max() < 2e-2 * scale
    assert np.linalg.norm(x) <= np.linalg.norm(x_cv) + 1e-3


def test_determinant_matches_cv2(oracle_lib):
    rng = np.random.RandomState(2)
    for _ in range(20):
        A = _spd6(rng, 1e2)
        d_cv = cv2.determinant(A)
        d = oracle_lib.det6(A)
        assert abs(d - d_cv) <= 1e-4 * abs(d_cv)
    assert oracle_lib.det6(np.zeros((6, 6), np.float32)) == 0.0
    assert cv2.determinant(np.zeros((6, 6), np.float32)) == 0.0


def test_rodrigues_matches_cv2(oracle_lib):
    rng = np.random.RandomState(4)
    for s in (1e-4, 1e-2, 0.5, 2.5):
        for _ in range(5):
            rv = (rng.randn(3) * s).astype(np.float32)
            t = rng.randn(3).astype(np.float32)
            T = oracle_lib.rodrigues(rv, t)
            R_cv, _ = cv2.Rodrigues(rv.astype(np.float64))
            assert np.abs(T[:3, :3] - R_cv).max() < 2e-7
            assert np.array_equal(T[:3, 3], t) and np.array_equal(T[3], [0, 0, 0, 1])
    assert np.array_equal(oracle_lib.rodrigues([0, 0, 0], [1, 2, 3])[:3, :3], np.eye(3, dtype=np.float32))


def test_pose_inverse_and_product_match_cv2(oracle_lib):
    rng = np.random.RandomState(5)
    for _ in range(20):
        R, _ = cv2.Rodrigues(rng.randn(3) * 0.3)
        P = np.eye(4, dtype=np.float32); P[:3, :3] = R; P[:3, 3] = rng.randn(3)
        ok, inv_cv = cv2.invert(P, flags=cv2.DECOMP_SVD)
        inv = oracle_lib.pose_inv(P)
        assert np.abs(inv - inv_cv).max() < 2e-6
        Q = np.eye(4, dtype=np.float32); Q[:3, :3] = cv2.Rodrigues(rng.randn(3) * 0.2)[0]; Q[:3, 3] = rng.randn(3)
        prod = oracle_lib.pose_mul(P, Q)
        assert np.abs(prod - (P.astype(np.float64) @ Q.astype(np.float64))).max() < 1e-6


def test_unpack_and_update_matches_cv2_chain(oracle_lib):
    """StreamHelper::get unpack order A00..A05,b0,A11..b5 + det test + solve + Tinc*affine (projective_icp.cpp:43-62,197-209)"""
    rng = np.random.RandomState(6)
    A = _spd6(rng, 1e2); A = ((A + A.T) / 2).astype(np.float32)
    b = (rng.randn(6) * 0.5).astype(np.float32)
    v27 = []
    for i in range(6):
        for j in range(i, 7):
            v27.append(b[i] if j == 6 else A[i, j])
    aff0 = np.eye(4, dtype=np.float32); aff0[:3, 3] = [0.01, -0.02, 0.03]
    ok, aff = oracle_lib.icp_solve_update(np.array(v27, np.float32), aff0)
    assert ok
    _, r = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)
    r = r.ravel()
    T = np.eye(4); T[:3, :3] = cv2.Rodrigues(r[:3].astype(np.float64))[0]; T[:3, 3] = r[3:]
    assert np.abs(aff - (T @ aff0.astype(np.float64))).max() < 1e-5
    ok0, _ = oracle_lib.icp_solve_update(np.zeros(27, np.float32), aff0)
    assert not ok0   # A = 0 -> |det| < 1e-15 -> tracking failure
