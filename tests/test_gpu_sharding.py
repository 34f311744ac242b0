"""GPU parity — the scene sharded over N contexts against ONE context holding the whole scene.

The N ranks are emulated in one process on one GPU: N contexts on the same stream, attached to one another by plain
device pointers (exactly what the multi-process path does with CUDA IPC pointers); stream order plays the role of the
cross-GPU barrier between the three frame stages.  The replicated index + owner-side payload + peer-memory raycast
must reproduce the single-context result bit for bit: poses, allocated set, every voxel, every ray, the visible set."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_shards(gpu, n, **kw):
    c0 = gpu.Context(shard_rank=0, shard_count=n, **kw)
    ctxs = [c0] + [gpu.Context(stream=c0.stream(), shard_rank=r, shard_count=n, **kw) for r in range(1, n)]
    ptrs = [c.shard_local_ptrs() for c in ctxs]
    for c in ctxs:
        for r, p in enumerate(ptrs):
            c.shard_attach(r, p)
    return ctxs


def sharded_frame(ctxs, dev_frame):
    for c in ctxs:
        c.frame_begin(dev_frame)
    for c in ctxs:
        c.frame_raycast()
    return [c.frame_end() for c in ctxs]


@pytest.mark.parametrize("n", [2, 3])
def test_sharded_scene_equals_single_context(gpu, s1_frames, n):
    depth, _, _ = s1_frames
    single = gpu.Context(corrected_mode=1)
    ctxs = make_shards(gpu, n, corrected_mode=1)
    try:
        frames = 6
        for i in range(frames):
            buf = ctxs[0].upload(depth[i], "frame")
            oks = sharded_frame(ctxs, buf)
            ok1 = single.process_frame(depth[i])
            assert all(o == ok1 for o in oks), (i, oks, ok1)
            p1 = single.pose()
            for c in ctxs:
                assert np.array_equal(c.pose().view(np.uint32), p1.view(np.uint32)), f"frame {i}: pose differs"
            assert sum(c.voxel_updates() for c in ctxs) == single.voxel_updates(), i
            r1 = single.raycast_result()
            if i > 0:
                for c in ctxs:
                    assert np.array_equal(c.raycast_result().view(np.uint32), r1.view(np.uint32)), f"frame {i}: raycast differs"
        # index replicated, payload partitioned
        t1 = single.table()
        set1 = gpu.allocated_set(t1)
        vis1 = gpu.visible_set(t1, single.visible_ids())
        b1 = single.blocks_by_pos()
        seen = {}
        for c in ctxs:
            t = c.table()
            assert gpu.allocated_set(t) == set1
            assert gpu.visible_set(t, c.visible_ids()) == vis1
            for pos, vox in c.blocks_by_pos().items():
                assert pos not in seen, "a block has two owners"
                seen[pos] = vox
        assert set(seen) == set(b1)
        for pos, vox in b1.items():
            assert np.array_equal(vox["sdf"], seen[pos]["sdf"]) and np.array_equal(vox["w"], seen[pos]["w"]), pos
        counts = [len(c.blocks_by_pos()) for c in ctxs]
        assert min(counts) > 0.6 * len(b1) / n, counts     # the owner mix balances the ranks
        # the viewer on a sharded rank reads foreign voxels through peer memory as well
        assert np.array_equal(ctxs[n - 1].render_image(), single.render_image())
    finally:
        for c in reversed(ctxs):   # rank 0 owns the shared stream
            c.close()
        single.close()


def test_sharded_context_refuses_the_unsharded_entry(gpu, s1_frames):
    depth, _, _ = s1_frames
    c = gpu.Context(shard_rank=1, shard_count=2)
    try:
        with pytest.raises(gpu.TfbError):
            c.process_frame(depth[0])
        buf = c.upload(depth[0], "frame")
        c.frame_begin(buf)
        c.frame_raycast()
        assert c.frame_end()
        c.frame_begin(c.upload(depth[1], "frame"))
        with pytest.raises(gpu.TfbError):      # peers were never attached
            c.frame_raycast()
    finally:
        c.close()


def test_sharded_scene_outside_the_directory_window(gpu, s1_frames):
    """the same at 1.5 mm voxels, where half the scene lies outside the block directory's window and is served by the hash —
    local blocks, this frame's copies of foreign blocks and the owner's pool alike"""
    depth, _, _ = s1_frames
    kw = dict(corrected_mode=1, voxel_size=0.0015, mu=0.006, num_blocks=1 << 18, num_buckets=1 << 21, excess_size=1 << 18)
    single = gpu.Context(**kw)
    ctxs = make_shards(gpu, 2, **kw)
    try:
        for i in range(3):
            buf = ctxs[0].upload(depth[i], "frame")
            oks = sharded_frame(ctxs, buf)
            ok1 = single.process_frame(depth[i])
            assert all(o == ok1 for o in oks), (i, oks, ok1)
            for c in ctxs:
                assert np.array_equal(c.pose().view(np.uint32), single.pose().view(np.uint32)), f"frame {i}: pose differs"
            assert sum(c.voxel_updates() for c in ctxs) == single.voxel_updates(), i
            if i > 0:
                r1 = single.raycast_result()
                for c in ctxs:
                    assert np.array_equal(c.raycast_result().view(np.uint32), r1.view(np.uint32)), f"frame {i}: raycast differs"
        t1 = single.table()
        for c in ctxs:
            assert gpu.visible_set(c.table(), c.visible_ids()) == gpu.visible_set(t1, single.visible_ids())
    finally:
        for c in reversed(ctxs):
            c.close()
        single.close()
