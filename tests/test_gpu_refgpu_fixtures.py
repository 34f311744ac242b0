"""GPU parity against the REFERENCE'S OWN DEVICE OUTPUT.

tests/golden/refgpu_*.npz were written by tests/golden/make_refgpu_golden.py, which runs the reference library's GPU
code (patched to compile: baseline/ref_gpu/patch_ref.py) on a B200 with the reference's nvcc flags
(--ftz=true --prec-div=false --prec-sqrt=false, fused multiply-adds on).  These tests put the CUDA path of this repo,
through the C ABI, on the same stored inputs.  They pin the stages whose arithmetic only exists on the device
(__expf, rsqrt, __fdividef) and add device-side evidence for the scene stages.

Bars (each measured on the B200 and written where it is asserted):
  integer / index work (dists, truncation, depth pyramid, vertex maps, resize, block sets, visible sets) .... bit-exact
  bilateral ................................. +-1 mm on < 1e-4 of the pixels (ex2.approx vs expf on a rounding tie)
  normal maps ............................... 2e-5 absolute (rsqrt.approx + FMA contraction vs IEEE 1/sqrt, no FMA)
  ICP 27-vector ............................. 1e-6 of the largest term (summation order), 2e-3 at the identity knife edge
  estimated transform ....................... 1e-4 m / 1e-4 rad (north-star pose tolerance); measured ~1e-7
  TSDF ...................................... 1e-3 of truncation (north-star) = 32 LSB, except voxels ON the boundary
                                              eta == -mu, where the reference's fused multiply-add decides (counted)
"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPORT = {}


def sha(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        a = np.where(np.isnan(a), np.float32(np.nan), a).astype(np.float32)
    return hashlib.sha256(a.tobytes()).hexdigest()


def rot_angle(Ra, Rb):
    D = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.arcsin(min(1.0, np.linalg.norm(w))))


def _load(name):
    p = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(p):
        pytest.fail(f"{p} is missing: regenerate with tests/golden/make_refgpu_golden.py on a GPU box")
    return np.load(p)


@pytest.fixture(scope="module")
def st():
    return _load("refgpu_stages_640x480")


@pytest.fixture(scope="module")
def sc():
    return _load("refgpu_scene_640x480")


@pytest.fixture(scope="module")
def ctx(gpu):
    c = gpu.Context()
    yield c
    c.close()


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    out = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "refgpu_fixture_parity.json"), "w") as f:
            json.dump(REPORT, f, indent=1)
    except OSError:
        pass


def _nan_close(a, b, tol, what):
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN masks differ at {int((na != nb).sum())} elements"
    d = np.abs(a[~na].astype(np.float64) - b[~na].astype(np.float64))
    REPORT[what] = {"max_abs": float(d.max()), "bit_equal_frac": float((a[~na].view(np.uint32) == b[~na].view(np.uint32)).mean())}
    assert d.max() <= tol, (what, d.max())


# ---- a1 .. a5 -------------------------------------------------------------------------------------------------
def test_a1_compute_dists_bit_exact(ctx, st):
    g = ctx.compute_dists(st["depth_in"])
    assert sha(g) == str(st["dists_sha"])
    assert np.array_equal(g[::8, ::8], st["dists_sample"])


def test_a2_bilateral_vs_reference_device(ctx, st):
    g = ctx.bilateral(st["depth_in"]).astype(np.int32)
    r = st["bilateral"].astype(np.int32)
    diff = np.abs(g - r)
    REPORT["bilateral"] = {"pixels_differing": int((diff > 0).sum()), "max_mm": int(diff.max())}
    assert diff.max() <= 1
    assert (diff > 0).mean() < 1e-4


def test_a3_truncate_bit_exact(ctx, st):
    assert np.array_equal(ctx.truncate_depth(st["bilateral"], 2.0), st["depth0"])


def test_a4_depth_pyramid_bit_exact(ctx, st):
    assert np.array_equal(ctx.depth_pyr(st["depth0"]), st["depth1"])
    assert np.array_equal(ctx.depth_pyr(st["depth1"]), st["depth2"])


def test_a5_vertex_maps_bit_exact_normal_maps_close(ctx, st):
    intr = st["intr"]
    for l in range(3):
        li = tuple(np.float32(v) / np.float32(1 << l) for v in intr)
        gp, gn = ctx.points_normals(st[f"depth{l}"], li)
        assert sha(gp) == str(st[f"points{l}_sha"]), f"vertex map level {l}"
        ref_n = {0: (st["normals0_s5"], 5), 1: (st["normals1_s3"], 3), 2: (st["normals2"], 1)}[l]
        _nan_close(gn[::ref_n[1], ::ref_n[1]], ref_n[0], 2e-5, f"normals_L{l}")
    assert np.array_equal(np.isnan(ctx.points_normals(st["depth2"], tuple(np.float32(v) / 4 for v in intr))[0]), np.isnan(st["points2"]))


def test_a17_resize_bit_exact(ctx, st):
    gp, gn = ctx.resize_points_normals(st["points2"], st["normals2"])
    assert sha(gp) == sha(st["resize_points"])
    assert sha(gn) == sha(st["resize_normals"])


# ---- a7 / a8 / a6 -----------------------------------------------------------------------------------------------
def test_a7_a8_icp_sums_from_the_reference_maps(ctx, st):
    """level-2 maps stored in the fixture: both sides reduce the SAME maps, only the reduction differs"""
    li = tuple(np.float32(v) / np.float32(4) for v in st["intr"])
    args = (st["icp_curr_points2"], st["icp_curr_normals2"], st["icp_model_points2"], st["icp_model_normals2"])
    for name, tol in (("small", 1e-6), ("orbit3", 1e-6), ("identity", 2e-3)):
        r27 = st[f"icp27_L2_{name}"]
        g27 = ctx.icp_reduce(li, st[f"icp_aff_{name}"], *args)
        rel = float(np.abs(g27 - r27).max() / np.abs(r27).max())
        REPORT[f"icp27_L2_{name}"] = rel
        assert rel <= tol, (name, rel)
    assert not st["icp27_L2_far"].any()
    assert not ctx.icp_reduce(li, st["icp_aff_far"], *args).any()


def test_a7_a8_icp_sums_end_to_end(gpu, st):
    """levels 0..2 from the stored depth frames through OUR preprocessing: image stages + reduction against the reference's"""
    c = gpu.Context()
    try:
        c.preprocess(st["icp_depth_model"])
        model = [(c.level(1, l), c.level(2, l)) for l in range(3)]
        c.preprocess(st["icp_depth_curr"])
        curr = [(c.level(1, l), c.level(2, l)) for l in range(3)]
        for l in range(3):
            li = tuple(np.float32(v) / np.float32(1 << l) for v in st["intr"])
            for name, tol in (("small", 2e-5), ("orbit3", 2e-5)):
                r27 = st[f"icp27_L{l}_{name}"]
                g27 = c.icp_reduce(li, st[f"icp_aff_{name}"], curr[l][0], curr[l][1], model[l][0], model[l][1])
                rel = float(np.abs(g27 - r27).max() / np.abs(r27).max())
                REPORT[f"icp27_e2e_L{l}_{name}"] = rel
                assert rel <= tol, (l, name, rel)
    finally:
        c.close()


def test_a6_estimate_transform_vs_reference_device(gpu, st):
    for k in (1, 3, 6):
        c = gpu.Context()
        try:
            c.preprocess(st["icp_depth_model"])
            for l in range(3):      # model maps <- the maps of the model frame
                c.set_level(3, l, c.level(1, l)); c.set_level(4, l, c.level(2, l))
            c.preprocess(st[f"est_depth_{k}"])
            ok, a = c.estimate_transform()
        finally:
            c.close()
        r = st[f"est_affine_{k}"]
        assert ok == bool(st[f"est_ok_{k}"])
        dt = float(np.abs(a[:3, 3] - r[:3, 3]).max()); dr = rot_angle(a[:3, :3], r[:3, :3])
        REPORT[f"estimate_0_{k}"] = {"dt_m": dt, "dr_rad": dr}
        assert dt < 1e-4 and dr < 1e-4, (k, dt, dr)
        assert np.abs(r[:3, 3]).max() > 1e-3     # a real motion was estimated


# ---- a9 .. a16 with injected poses ---------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=[0, 1], ids=["device_arith", "ieee_arith"])
def scene_run(gpu, sc, request):
    """ieee_arith=0 (default): TSDF integration in the arithmetic of the reference's GPU build; 1: IEEE (the CPU oracle's)"""
    g = gpu.Context(ieee_arith=request.param)
    sets = {}
    da, db = g.compute_dists(sc["depth_a"]), g.compute_dists(sc["depth_b"])
    wa, wb = sc["pose_a_w2c"], sc["pose_b_w2c"]
    g.allocate(wa, da); g.integrate(wa, da)
    sets["pass1"] = gpu.allocated_set(g.table())
    g.allocate(wa, da); g.integrate(wa, da)
    sets["pass2"] = gpu.allocated_set(g.table())
    g.allocate(wb, db); g.integrate(wb, db)
    sets["final"] = gpu.allocated_set(g.table())
    sets["visible"] = gpu.visible_set(g.table(), g.visible_ids())
    blocks = g.blocks_by_pos()
    g.expected_depths(wb)
    minmax = g.minmax()
    maps = g.icp_maps(sc["pose_b"])
    ray = g.raycast_result()
    lv = {l: (g.level(3, l), g.level(4, l)) for l in range(3)}
    yield {"sets": sets, "blocks": blocks, "minmax": minmax, "maps": maps, "ray": ray, "levels": lv, "ctx": g, "ieee": request.param}
    g.close()


def _as_set(pos):
    return set(map(tuple, np.asarray(pos).tolist()))


def test_a10_a11_block_set_vs_reference_device(scene_run, sc):
    """The reference's allocation loses same-hash collisions to a race (plain stores, last writer wins; the loser —
    or, when three new blocks meet in one slot, the losers — are allocated one frame later each), so its set after the first
    pass is one of several legal outcomes.  Every block either side holds must be a block the frame requests, and once the
    collisions have drained (the third allocation here) the sets must be identical, block for block."""
    s = scene_run["sets"]
    ref1, ref2, ref3 = _as_set(sc["blocks_pass1"]), _as_set(sc["blocks_pass2"]), _as_set(sc["blocks"])
    REPORT["blocks"] = {"pass1_ours": len(s["pass1"]), "pass1_ref": len(ref1), "pass1_symdiff": len(s["pass1"] ^ ref1),
                        "pass2_symdiff": len(s["pass2"] ^ ref2), "final_symdiff": len(s["final"] ^ ref3), "final": len(ref3)}
    assert s["pass1"] <= s["pass2"] <= s["final"] and ref1 <= ref2 <= ref3
    assert len(s["pass1"] ^ ref1) <= 0.05 * len(ref1)      # ~n^2/2m contested slots of ~1 800 new blocks (measured: 45)
    assert len(s["pass2"] ^ ref2) <= 4                     # only third-order collisions are left (measured: 2)
    assert s["final"] == ref3


def test_a12_visible_set_vs_reference_device(scene_run, sc):
    assert scene_run["sets"]["visible"] == _as_set(sc["visible_blocks"])


def test_a13_voxels_vs_reference_device(scene_run, sc):
    """Blocks whose weights agree everywhere were integrated by the same frames on both sides (the rest lost the reference's
    allocation race in pass 1 and have one integration less there).  On those blocks:
      * default arithmetic (the reference's device build re-issued instruction for instruction): EVERY voxel bit-identical;
      * IEEE arithmetic (bit-identical to a host compile instead): north-star tolerance, 1e-3 of truncation = 32 LSB of 32767
        (measured on the B200: 1.3 % of the voxels differ, by a few LSB; 2 of 974 336 by more than 32)."""
    ours = scene_run["blocks"]
    n_blocks = n_same_w = 0
    n_vox = n_diff = n_over = 0
    max_lsb = 0
    for pos, sdf, w in zip(sc["blocks"], sc["sdf"], sc["w"]):
        b = ours[tuple(int(v) for v in pos)]
        n_blocks += 1
        if not np.array_equal(b["w"], w):
            continue
        n_same_w += 1
        d = np.abs(b["sdf"].astype(np.int32) - sdf.astype(np.int32))
        n_vox += 512; n_diff += int((d > 0).sum()); n_over += int((d > 32).sum()); max_lsb = max(max_lsb, int(d.max()))
    REPORT["voxels_" + ("ieee" if scene_run["ieee"] else "device")] = {
        "blocks": n_blocks, "blocks_same_weights": n_same_w, "voxels": n_vox, "differing": n_diff, "beyond_1e-3_of_truncation": n_over, "max_lsb": max_lsb}
    assert n_same_w >= 0.9 * n_blocks
    if scene_run["ieee"]:
        assert n_over <= 1e-4 * n_vox
    else:
        assert n_diff == 0, f"{n_diff} of {n_vox} voxels differ from the reference's GPU output (max {max_lsb} LSB)"


@pytest.mark.parametrize("ieee", [0, 1], ids=["device_arith", "ieee_arith"])
def test_a13_identity_pose_boundary_voxels_are_counted(gpu, sc, ieee):
    """Frame 0 (pose = identity): voxel planes sit exactly on eta == -mu for depths that are multiples of 5 mm; the side the
    reference's device code lands on is decided by its fused multiply-add.  Everything off that boundary must agree."""
    g = gpu.Context(ieee_arith=ieee)
    try:
        d = g.compute_dists(sc["identity_depth"])
        eye = np.eye(4, dtype=np.float32)
        for _ in range(3):
            g.allocate(eye, d); g.integrate(eye, d)
        ours = g.blocks_by_pos()
    finally:
        g.close()
    ref_keys = _as_set(sc["identity_blocks"])
    assert ref_keys <= set(ours.keys()) and len(set(ours.keys()) - ref_keys) <= 4     # ours ran one more pass (collisions drained)
    n = nd = n_big = 0
    for pos, sdf, w in zip(sc["identity_blocks"], sc["identity_sdf"], sc["identity_w"]):
        b = ours[tuple(int(v) for v in pos)]
        touched_same = (b["w"] > 0) == (w > 0)
        # ours ran three passes, the reference two: running averages of one and the same value, so sdf is comparable where both touched
        both = (b["w"] > 0) & (w > 0)
        dd = np.abs(b["sdf"].astype(np.int32) - sdf.astype(np.int32))
        n += 512; nd += int((~touched_same).sum()); n_big += int((both & (dd > 32)).sum())
    REPORT["identity_pose_" + ("ieee" if ieee else "device")] = {"voxels": n, "on_boundary_decided_differently": nd, "other_beyond_tolerance": n_big}
    assert n_big <= 1e-4 * n
    if not ieee:
        assert nd == 0      # the reference's own arithmetic: the boundary voxels fall on the same side


def test_a14_expected_depth_vs_reference_device(scene_run, sc):
    g, r = scene_run["minmax"], sc["range_image"]
    d = np.abs(g.astype(np.float64) - r.astype(np.float64))
    REPORT["range_image"] = {"max_abs": float(d.max()), "bit_equal_frac": float((g.view(np.uint32) == r.view(np.uint32)).mean())}
    assert d.max() <= 1e-3 * max(1.0, float(np.abs(r[r < 1e5]).max()))


def test_a15_raycast_vs_reference_device(scene_run, sc):
    g = scene_run["ray"]
    hit_g = np.packbits(g[..., 3] > 0)
    hits_differ = int(np.unpackbits(hit_g ^ sc["raycast_hit_mask"]).sum())
    gs, rs = g[::4, ::4], sc["raycast_s4"]
    both = (gs[..., 3] > 0) & (rs[..., 3] > 0)
    d = np.abs(gs[both][:, :3].astype(np.float64) - rs[both][:, :3].astype(np.float64))   # voxel units
    REPORT["raycast"] = {"hit_mask_differs": hits_differ, "of": int(g.shape[0] * g.shape[1]), "max_voxels": float(d.max()),
                         "p999_voxels": float(np.quantile(d, 0.999)), "bit_equal": bool(sha(g) == str(sc["raycast_sha"]))}
    assert hits_differ <= 2e-3 * g.shape[0] * g.shape[1]
    assert np.quantile(d, 0.999) <= 0.05                       # 5 % of a voxel = 0.25 mm at 5 mm voxels


def test_a16_a17_model_maps_vs_reference_device(scene_run, sc):
    ctx = scene_run["ctx"]
    p0, n0 = scene_run["maps"]
    p1, n1 = ctx.resize_points_normals(p0, n0)
    p2, n2 = ctx.resize_points_normals(p1, n1)
    for name, (gp, gn), (rp, rn) in (("L0_s5", (p0[::5, ::5], n0[::5, ::5]), (sc["model_points0_s5"], sc["model_normals0_s5"])),
                                     ("L2", (p2, n2), (sc["model_points2"], sc["model_normals2"]))):
        both = ~np.isnan(gp[..., 0]) & ~np.isnan(rp[..., 0])
        mism = int((np.isnan(gp[..., 0]) != np.isnan(rp[..., 0])).sum())
        dp = np.abs(gp[both][:, :3].astype(np.float64) - rp[both][:, :3].astype(np.float64))
        dn = np.abs(gn[both][:, :3].astype(np.float64) - rn[both][:, :3].astype(np.float64))
        REPORT["model_maps_" + name] = {"validity_differs": mism, "of": int(both.size), "points_p999_m": float(np.quantile(dp, 0.999)),
                                        "normals_p999": float(np.quantile(dn, 0.999)), "points_max_m": float(dp.max()), "normals_max": float(dn.max())}
        assert mism <= 0.01 * both.size
        assert np.quantile(dp, 0.999) <= 1e-3      # 1 mm; typical agreement is ~1e-6 m (see the report)
