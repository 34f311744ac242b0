"""GPU parity — voxel-hash allocation, visible list, TSDF integration, expected depth, raycast and model
maps against the oracle with INJECTED poses, so every stage is decoupled from ICP.  These stages are integer
/ IEEE-exact work: the bar is bit-exact (block sets, visible sets, every voxel, every ray)."""
import numpy as np
import pytest

from conftest import same_bits_nan

pytestmark = pytest.mark.gpu


def _run_both(gpu, tfo, depth, poses, n, **kw):
    """alloc+integrate+expected-depth+icp-maps for n frames with ground-truth poses on both sides"""
    L = tfo.Lib("port")
    o = tfo.Oracle(lib=L, **kw)
    g = gpu.Context(ieee_arith=1, **kw)   # IEEE arithmetic: the CPU oracle is a host compile of the reference
    out = []
    for i in range(n):
        dists = L.compute_dists(depth[i], o.params.depth_cutoff_mm)
        c2w = poses[i].astype(np.float32)
        w2c = L.pose_inv(c2w)
        o.allocate(w2c, dists); g.allocate(w2c, dists)
        o.integrate(w2c, dists); g.integrate(w2c, dists)
        o.expected_depths(w2c); g.expected_depths(w2c)
        om = o.icp_maps(c2w); gm = g.icp_maps(c2w)
        out.append((om, gm))
    return o, g, out


@pytest.fixture(scope="module")
def both(gpu, s1_frames):
    from oracle import tfo
    depth, poses, _ = s1_frames
    o, g, maps = _run_both(gpu, tfo, depth, poses, 5)
    yield o, g, maps, tfo
    g.close(); o.close()


def test_allocated_block_set_bit_exact(both):
    o, g, _, tfo = both
    so, sg = tfo.allocated_set(o.table()), tfo.allocated_set(g.table())
    assert len(so) > 1000
    assert so == sg


def test_counters_match(both):
    o, g, _, _ = both
    co, cg = o.counters(), g.counters()
    for k in ("n_visible", "last_free_block", "last_free_excess", "n_allocated"):
        assert co[k] == cg[k], (k, co[k], cg[k])
    assert o.voxel_updates() == g.voxel_updates() > 0


def test_visible_list_bit_exact_as_set(both):
    o, g, _, tfo = both
    vo = tfo.visible_set(o.table(), o.visible_ids())
    vg = tfo.visible_set(g.table(), g.visible_ids())
    assert vo == vg
    ids = g.visible_ids()
    assert len(ids) == len(set(ids.tolist())), "duplicate entries in the visible list"


def test_hash_chain_structure_consistent(both):
    """every allocated block is reachable from its bucket through the excess chain, exactly once"""
    _, g, _, _ = both
    t = g.table()
    nb = g.params.num_buckets
    alloc = np.nonzero(t["ptr"] >= 0)[0]
    seen = set()
    for s in alloc:
        pos = tuple(int(v) for v in t["pos"][s])
        assert pos not in seen
        seen.add(pos)
        h = ((np.uint32(pos[0] & 0xffffffff) * np.uint32(73856093)) ^ (np.uint32(pos[1] & 0xffffffff) * np.uint32(19349669))
             ^ (np.uint32(pos[2] & 0xffffffff) * np.uint32(83492791))) & np.uint32(nb - 1)
        cur = int(h)
        for _ in range(64):
            if cur == s:
                break
            assert t["offset"][cur] >= 1, "chain broken"
            cur = nb + int(t["offset"][cur]) - 1
        assert cur == s
    ptrs = t["ptr"][alloc]
    assert len(set(ptrs.tolist())) == len(ptrs), "two blocks share a pool slot"


def test_every_voxel_bit_exact(both):
    o, g, _, _ = both
    bo, bg = o.blocks_by_pos(), g.blocks_by_pos()
    assert bo.keys() == bg.keys()
    touched = 0
    for k in bo:
        a, b = bo[k], bg[k]
        assert np.array_equal(a["sdf"], b["sdf"]), k
        assert np.array_equal(a["w"], b["w"]), k
        touched += int((a["w"] > 0).sum())
    assert touched > 100000


def test_expected_depth_bit_exact(both):
    o, g, _, _ = both
    mo = np.ascontiguousarray(o.minmax()[: o.rows // 8, : o.cols // 8])
    assert np.array_equal(mo.view(np.uint32), g.minmax().view(np.uint32))


def test_raycast_bit_exact(both):
    o, g, _, _ = both
    ro, rg = o.raycast_result(), g.raycast_result()
    hit_o, hit_g = ro[..., 3] > 0, rg[..., 3] > 0
    assert hit_o.sum() > 50000
    assert np.array_equal(hit_o, hit_g)
    assert np.array_equal(ro.view(np.uint32), rg.view(np.uint32))


def test_icp_maps_bit_exact(both):
    _, _, maps, _ = both
    for (om, gm) in maps:
        for a, b in zip(om, gm):
            assert same_bits_nan(a, b).all()
            va, vb = ~np.isnan(a[..., 0]), ~np.isnan(b[..., 0])
            assert np.array_equal(a[va].view(np.uint32), b[vb].view(np.uint32))


def test_vis_type_state_matches(both):
    """visibility state incl. the raycast feedback and the slot-0 quirk"""
    o, g, _, _ = both
    assert np.array_equal(o.vis_type() > 0, g.vis_type() > 0)
    assert g.vis_type()[0] == 1


@pytest.mark.parametrize("voxel,mu", [(0.008, 0.02), (0.010, 0.02), (0.004, 0.016), (0.005, 0.04)])
def test_other_voxel_sizes_incl_knife_edges(gpu, s1_frames, voxel, mu):
    """8 mm = BASELINE config 1; 10 mm makes noSteps = ceil(2*0.5) sit on the integer"""
    from oracle import tfo
    depth, poses, _ = s1_frames
    o, g, _ = _run_both(gpu, tfo, depth, poses, 3, voxel_size=voxel, mu=mu)
    try:
        assert tfo.allocated_set(o.table()) == tfo.allocated_set(g.table())
        assert o.counters()["n_visible"] == g.counters()["n_visible"]
        bo, bg = o.blocks_by_pos(), g.blocks_by_pos()
        for k in list(bo)[::7]:
            assert np.array_equal(bo[k]["sdf"], bg[k]["sdf"]) and np.array_equal(bo[k]["w"], bg[k]["w"])
        assert np.array_equal(o.raycast_result().view(np.uint32), g.raycast_result().view(np.uint32))
    finally:
        g.close(); o.close()


def test_empty_frame_allocates_nothing(gpu):
    g = gpu.Context(ieee_arith=1)
    try:
        z = np.full((480, 640), -1.0, np.float32)
        g.allocate(np.eye(4, dtype=np.float32), z)
        g.integrate(np.eye(4, dtype=np.float32), z)
        c = g.counters()
        assert c["n_visible"] == 0 and c["n_allocated"] == 0 and g.voxel_updates() == 0
    finally:
        g.close()


def test_pool_exhaustion_restores_counters(gpu, s1_frames):
    """out of blocks: silently skip and roll the counters back (SceneReconstructionEngine_host.cu:374-381)"""
    from oracle import tfo
    depth, poses, _ = s1_frames
    L = tfo.Lib("port")
    g = gpu.Context(num_blocks=500, ieee_arith=1)
    try:
        dists = L.compute_dists(depth[0])
        for _ in range(2):
            g.allocate(np.eye(4, dtype=np.float32), dists)
        c = g.counters()
        assert c["n_allocated"] == 500 and c["last_free_block"] == -1
        t = g.table()
        assert (t["ptr"] >= 0).sum() == 500
    finally:
        g.close()


def test_viewer_render_bit_exact(both):
    """TopFu::renderImage (RENDER_SHADED_GREYSCALE from a new raycast): SDF-gradient normals, 32 voxel reads per pixel"""
    o, g, _, _ = both
    pose = o.L.pose_inv(np.eye(4, dtype=np.float32))
    vis_before = g.vis_type().copy()
    io, ig = o.render_image(pose), g.render_image(pose)
    assert (io[..., 0] > 0).mean() > 0.2
    assert np.array_equal(io, ig)
    assert np.array_equal(vis_before, g.vis_type()), "renderImage must not touch the visible set (updateVisibleList=false)"


def test_ragged_depth_with_holes_and_noise_bit_exact(gpu, s1_frames):
    """30 % of the pixels knocked out at random, 1 mm Gaussian noise on the rest: the block set, every voxel and every ray must
    still match the oracle bit for bit (ragged segments, isolated pixels, blocks allocated from a single sample)"""
    from oracle import tfo
    depth, poses, _ = s1_frames
    rng = np.random.RandomState(5)
    noisy = []
    for i in range(3):
        d = depth[i].astype(np.float64)
        d = np.where(d > 0, np.rint(d + rng.normal(0, 1.0, d.shape)), 0)
        d[rng.rand(*d.shape) < 0.3] = 0
        noisy.append(np.clip(d, 0, 65535).astype(np.uint16))
    o, g, maps = _run_both(gpu, tfo, noisy, poses, 3)
    try:
        assert tfo.allocated_set(o.table()) == tfo.allocated_set(g.table())
        assert tfo.visible_set(o.table(), o.visible_ids()) == tfo.visible_set(g.table(), g.visible_ids())
        assert o.voxel_updates() == g.voxel_updates() > 0
        bo, bg = o.blocks_by_pos(), g.blocks_by_pos()
        for k in sorted(bo)[::3]:
            assert np.array_equal(bo[k]["sdf"], bg[k]["sdf"]) and np.array_equal(bo[k]["w"], bg[k]["w"]), k
        assert np.array_equal(o.raycast_result().view(np.uint32), g.raycast_result().view(np.uint32))
        (op, on), (gp, gn) = maps[-1]
        assert same_bits_nan(op, gp).all() and same_bits_nan(on, gn).all()
    finally:
        g.close(); o.close()
