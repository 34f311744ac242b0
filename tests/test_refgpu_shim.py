"""The patched reference GPU build (baseline/ref_gpu) replaces OpenCV — absent from this image — by a small shim
(baseline/ref_gpu/shim/opencv2).  The reference calls cv::determinant, cv::solve(DECOMP_SVD), Affine3f(rvec, t), the
Affine3f product and inverse on its hot path (projective_icp.cpp:197-209, topfu.cpp:243,281); these tests pin the shim's
versions against cv2 4.13, so that poses produced by baseline/_ref/libref_gpu*.so are the reference's poses.  CPU only; the
patch script's own checks (every edit must match exactly N times) run here too."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def L():
    from baseline.ref_gpu import refgpu
    if not refgpu.available():
        pytest.skip("baseline/_ref/libref_gpu_nodebug.so not built (needs /root/reference: make -C baseline/ref_gpu)")
    return refgpu.lib(), refgpu._p


def _spd6(rng, cond):
    q, _ = np.linalg.qr(rng.randn(6, 6))
    w = np.logspace(0, np.log10(cond), 6) * 50.0
    a = (q * w) @ q.T
    return ((a + a.T) / 2).astype(np.float32)


def test_shim_solve_and_determinant_match_cv2(L):
    lib, p = L
    rng = np.random.RandomState(3)
    for cond in (10, 1e3, 1e5):
        for _ in range(8):
            A = _spd6(rng, cond); b = (rng.randn(6) * 10).astype(np.float32)
            x = np.zeros(6, np.float32)
            lib.refgpu_cv_solve6(p(A), p(b), p(x))
            _, x_cv = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_SVD)
            ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
            scale = np.abs(ref).max()
            assert np.abs(x - ref).max() <= 2e-6 * scale
            assert np.abs(x - x_cv.ravel()).max() <= max(5e-4 * cond / 1e3, 5e-5) * scale
            d = lib.refgpu_cv_determinant6(p(A))
            assert abs(d - cv2.determinant(A)) <= 1e-4 * abs(cv2.determinant(A))


def test_shim_affine_matches_cv2_rodrigues_and_matrix_algebra(L):
    lib, p = L
    rng = np.random.RandomState(4)
    for _ in range(20):
        rv = (rng.randn(3) * 0.05).astype(np.float32); t = rng.randn(3).astype(np.float32)
        m = np.zeros((4, 4), np.float32)
        lib.refgpu_cv_affine(p(rv), p(t), p(m))
        R_cv, _ = cv2.Rodrigues(rv.astype(np.float64))
        assert np.abs(m[:3, :3] - R_cv).max() <= 1.2e-7
        assert np.array_equal(m[:3, 3], t) and np.array_equal(m[3], [0, 0, 0, 1])
        rv2 = (rng.randn(3) * 0.3).astype(np.float32); t2 = rng.randn(3).astype(np.float32)
        m2 = np.zeros((4, 4), np.float32); prod = np.zeros((4, 4), np.float32); inv = np.zeros((4, 4), np.float32)
        lib.refgpu_cv_affine(p(rv2), p(t2), p(m2))
        lib.refgpu_cv_affine_mul(p(m), p(m2), p(prod))
        assert np.abs(prod - m.astype(np.float64) @ m2.astype(np.float64)).max() <= 5e-7
        lib.refgpu_cv_affine_inv(p(prod), p(inv))
        ok, inv_cv = cv2.invert(prod, flags=cv2.DECOMP_SVD)
        assert np.abs(inv - inv_cv).max() <= 2e-6


def test_patch_script_applies_cleanly_when_the_reference_is_present(tmp_path):
    if not os.path.isdir("/root/reference/tfusion/src"):
        pytest.skip("reference tree absent (GPU box)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_gpu", "patch_ref.py"), "/root/reference", str(tmp_path / "gen")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = open("/root/reference/tfusion/src/topfu.cpp", encoding="latin-1").read()
    gen = open(tmp_path / "gen" / "src" / "topfu.cpp", encoding="latin-1").read()
    assert ref == gen                                    # the as-shipped variant is the file itself
    nodebug = open(tmp_path / "gen" / "src" / "topfu_nodebug.cpp", encoding="latin-1").read()
    assert len(ref.split("\n")) - len(nodebug.split("\n")) == 13
    assert "estimateTransform(affine,p.intr,curr_.points_pyr" in nodebug and "renderImage(image);" not in nodebug
