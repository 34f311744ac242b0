"""GPU — the step after the path: surface-point extraction from the voxel-block hash against a numpy
restatement on the very same voxels (bit-exact as a set), and scene save / load with tracking resumed on the restored model."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def numpy_points(blocks: dict, voxel_size: float) -> np.ndarray:
    """zero crossings of the TSDF along +x, +y, +z voxel edges (both voxels observed), linearly interpolated; float32 like the kernel"""
    vs = np.float32(voxel_size)
    out = []
    for (bx, by, bz), v in blocks.items():
        sdf = v["sdf"].reshape(8, 8, 8).astype(np.float32)      # [z][y][x]
        w = v["w"].reshape(8, 8, 8)
        f0 = sdf / np.float32(32767.0)
        for ax, (dz, dy, dx) in enumerate(((0, 0, 1), (0, 1, 0), (1, 0, 0))):
            nb = blocks.get((bx + dx, by + dy, bz + dz))
            # neighbour values: shifted inside the block, first slice of the +1 block across the face
            f1 = np.full((8, 8, 8), np.nan, np.float32)
            w1 = np.zeros((8, 8, 8), np.uint8)
            src = (slice(dz, None), slice(dy, None), slice(dx, None))
            dst = (slice(0, 8 - dz), slice(0, 8 - dy), slice(0, 8 - dx))
            f1[dst] = f0[src]
            w1[dst] = w[src]
            if nb is not None:
                nf = nb["sdf"].reshape(8, 8, 8).astype(np.float32) / np.float32(32767.0)
                nw = nb["w"].reshape(8, 8, 8)
                face = (slice(7, 8) if dz else slice(None), slice(7, 8) if dy else slice(None), slice(7, 8) if dx else slice(None))
                nface = (slice(0, 1) if dz else slice(None), slice(0, 1) if dy else slice(None), slice(0, 1) if dx else slice(None))
                f1[face] = nf[nface]
                w1[face] = nw[nface]
            with np.errstate(invalid="ignore"):
                hit = (w > 0) & (w1 > 0) & (((f0 > 0) & (f1 < 0)) | ((f0 < 0) & (f1 > 0)))
            z, y, x = np.nonzero(hit)
            if len(z) == 0:
                continue
            t = f0[hit] / (f0[hit] - f1[hit])
            g = np.stack([(bx * 8 + x).astype(np.float32), (by * 8 + y).astype(np.float32), (bz * 8 + z).astype(np.float32)], axis=1)
            g[:, ax] = g[:, ax] + t
            out.append(g * vs)
    p = np.concatenate(out) if out else np.zeros((0, 3), np.float32)
    return p.astype(np.float32)


def _sorted_rows(a):
    a = np.ascontiguousarray(a[:, :3]).view(np.uint32)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_extract_points_matches_numpy_restatement(gpu, s1_frames):
    depth, poses, _ = s1_frames
    g = gpu.Context(corrected_mode=1)
    try:
        for i in range(4):
            assert g.process_frame(depth[i])
        pts = g.extract_points()
        assert pts.shape[0] > 50000 and np.all(pts[:, 3] == 1.0)
        ref = numpy_points(g.blocks_by_pos(), 0.005)
        assert ref.shape[0] == pts.shape[0]
        assert np.array_equal(_sorted_rows(pts), _sorted_rows(ref))
        # the points are the scene: the sphere of radius 0.35 m at (0, 0, 1.2) is among them
        d = np.linalg.norm(pts[:, :3] - np.array([0, 0, 1.2], np.float32), axis=1)
        on_sphere = np.abs(d - 0.35) < 0.004
        assert on_sphere.sum() > 5000
    finally:
        g.close()


def test_scene_save_load_and_resume_tracking(gpu, s1_frames, tmp_path):
    depth, poses, _ = s1_frames
    a = gpu.Context(corrected_mode=1)
    b = gpu.Context(corrected_mode=1)
    try:
        for i in range(5):
            assert a.process_frame(depth[i])
        path = str(tmp_path / "scene.tfb")
        a.save_scene(path)
        b.load_scene(path)
        assert b.num_poses() == a.num_poses() and np.array_equal(a.pose(), b.pose())
        assert gpu.allocated_set(a.table()) == gpu.allocated_set(b.table())
        ba, bb = a.blocks_by_pos(), b.blocks_by_pos()
        for k in ba:
            assert np.array_equal(ba[k]["sdf"], bb[k]["sdf"]) and np.array_equal(ba[k]["w"], bb[k]["w"])
        assert np.array_equal(_sorted_rows(a.extract_points()), _sorted_rows(b.extract_points()))
        # tracking goes on against the restored model (the visible list is rebuilt, not restored: 1e-5 m, not bit-exact)
        for i in range(5, 8):
            assert a.process_frame(depth[i]) and b.process_frame(depth[i])
            assert np.abs(a.pose()[:3, 3] - b.pose()[:3, 3]).max() < 1e-5, i
        with pytest.raises(gpu.TfbError):
            gpu.Context(voxel_size=0.004).load_scene(path)     # written with another voxel size
    finally:
        a.close(); b.close()


def test_view_point_cloud_matches_the_reference_function(gpu, s1_frames):
    """tfb_render_point_cloud against the oracle's render_point_cloud, which follows the reference's own (dormant)
    renderPointCloud_device (include/tfusion/cuda/VisualisationHelper.hpp:150-198) with the reference's castRay and
    computeNormalAndAngle underneath (oracle/_ref build; bit-identical port otherwise).  Same scene on both sides (injected
    ground-truth poses, IEEE integration arithmetic = the oracle's), same view, with and without skipPoints: the clouds must be
    the same SET of points, bit for bit — the order is unspecified on both sides (CTA arrival order in the reference)."""
    from oracle import tfo
    depth, poses, _ = s1_frames
    L = tfo.Lib("ref" if tfo.have_ref() else "port")
    o = tfo.Oracle(lib=L)
    g = gpu.Context(ieee_arith=1)
    try:
        for i in range(3):
            dists = L.compute_dists(depth[i])
            w2c = L.pose_inv(poses[i].astype(np.float32))
            o.allocate(w2c, dists); g.allocate(w2c, dists)
            o.integrate(w2c, dists); g.integrate(w2c, dists)
            o.expected_depths(w2c); g.expected_depths(w2c)
        view = poses[2].astype(np.float32)
        for skip in (False, True):
            po = o.render_point_cloud(view, skip)
            pg = g.render_point_cloud(view, skip)
            assert po.shape[0] > (10000 if skip else 40000)
            assert pg.shape == po.shape
            assert np.all(pg[:, 3] == 1.0)
            assert np.array_equal(_sorted_rows(pg), _sorted_rows(po))
        # the view sees the sphere of radius 0.35 m at (0, 0, 1.2)
        d = np.linalg.norm(pg[:, :3] - np.array([0, 0, 1.2], np.float32), axis=1)
        assert (np.abs(d - 0.35) < 0.004).sum() > 2000
    finally:
        g.close(); o.close()
