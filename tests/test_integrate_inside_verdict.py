"""k_integrate's per-block verdict (topfusion_b200/csrc/tfb_scene.cu, block_projects_inside) is allowed to say "inside" only when
EVERY voxel of the block passes the reference's early returns (computeUpdatedVoxelDepthInfo, SceneReconstructionEngine.hpp:33-41:
z > 0, 1 <= u <= w - 2, 1 <= v <= h - 2) and sits farther than mu from the camera plane.  The verdict is restated here in fp32
numpy, operation for operation, and checked against the exact projection of all 512 voxels on random poses and blocks — the
property the GPU's bit-exact voxel tests rely on at the image border, pinned on the host."""
import numpy as np

W, H = 640, 480
FX, FY, CX, CY = 504.261, 503.905, 352.457, 272.202
f32 = np.float32


def verdict(M, bx, by, bz, voxel, mu):
    """fp32 restatement of block_projects_inside (M: 4x4 world -> camera, row-major here)"""
    h = f32(3.5) * f32(voxel)
    p = [f32(b * 8) * f32(voxel) + h for b in (bx, by, bz)]
    M = M.astype(np.float32)
    c = [M[r, 0] * p[0] + M[r, 1] * p[1] + M[r, 2] * p[2] + M[r, 3] for r in range(3)]
    a = [h * (abs(M[r, 0]) + abs(M[r, 1]) + abs(M[r, 2])) for r in range(3)]
    z_lo, z_hi = c[2] - a[2], c[2] + a[2]
    if not (z_lo > f32(1.01) * f32(mu) + f32(1e-3)):
        return False
    r_lo, r_hi = f32(1) / z_lo, f32(1) / z_hi
    x_lo, x_hi, y_lo, y_hi = c[0] - a[0], c[0] + a[0], c[1] - a[1], c[1] + a[1]
    u_min = f32(FX) * min(x_lo * r_lo, x_lo * r_hi) + f32(CX)
    u_max = f32(FX) * max(x_hi * r_lo, x_hi * r_hi) + f32(CX)
    v_min = f32(FY) * min(y_lo * r_lo, y_lo * r_hi) + f32(CY)
    v_max = f32(FY) * max(y_hi * r_lo, y_hi * r_hi) + f32(CY)
    return bool(u_min >= 2 and v_min >= 2 and u_max <= W - 3 and v_max <= H - 3)


def all_voxels_pass(M, bx, by, bz, voxel, mu):
    o = np.arange(8)
    g = np.stack(np.meshgrid(o + bx * 8, o + by * 8, o + bz * 8, indexing="ij"), -1).reshape(-1, 3).astype(np.float64) * voxel
    pc = g @ M[:3, :3].T + M[:3, 3]
    z = pc[:, 2]
    if not (z > mu).all():
        return False, 0.0
    u = FX * pc[:, 0] / z + CX
    v = FY * pc[:, 1] / z + CY
    slack = min(u.min() - 1, (W - 2) - u.max(), v.min() - 1, (H - 2) - v.max())
    return bool(slack >= 0), float(slack)


def random_pose(rng):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = rng.uniform(-np.pi, np.pi)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = rng.uniform(-1.5, 1.5, size=3)
    return M


def test_inside_verdict_is_conservative():
    rng = np.random.default_rng(20261018)
    said_inside = 0
    worst = np.inf
    for voxel, mu in ((0.005, 0.02), (0.002, 0.016), (0.01, 0.02), (0.008, 0.064)):
        for _ in range(120):
            M = random_pose(rng)
            Minv = np.linalg.inv(M)
            # blocks around points the camera actually looks at, border of the image included
            for _ in range(40):
                z = rng.uniform(0.05, 3.5)
                u, v = rng.uniform(-40, W + 40), rng.uniform(-40, H + 40)
                pw = Minv[:3, :3] @ np.array([(u - CX) / FX * z, (v - CY) / FY * z, z]) + Minv[:3, 3]
                b = np.floor(pw / (8 * voxel)).astype(int)
                if verdict(M, b[0], b[1], b[2], voxel, mu):
                    ok, slack = all_voxels_pass(M, b[0], b[1], b[2], voxel, mu)
                    assert ok, (voxel, mu, b, slack)
                    said_inside += 1
                    worst = min(worst, slack)
    assert said_inside > 3000          # the test is not vacuous: most in-image blocks take the short path
    assert worst > 0.5                 # and the verdict keeps about a pixel of margin, far above fp32 rounding (1e-4 px)


def test_inside_verdict_accepts_the_bulk_of_a_frame():
    """identity pose, blocks 1 m in front of the camera: everything but a ring along the image border takes the short path"""
    M = np.eye(4)
    voxel, mu = 0.005, 0.02
    took = total = 0
    for bx in range(-20, 21):
        for by in range(-16, 17):
            ok_all, _ = all_voxels_pass(M, bx, by, 25, voxel, mu)
            if ok_all:
                total += 1
                took += verdict(M, bx, by, 25, voxel, mu)
    assert total > 300 and took >= 0.8 * total
