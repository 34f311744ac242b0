"""Block streaming between the voxel pool and host memory (tfb_stream_out / tfb_stream_in): a scene whose blocks
are evicted when they leave the enlarged frustum and restored before the camera looks at them again is IDENTICAL to a scene that was
never streamed — poses per frame, block set, every voxel.  The reference carries the mechanism as dormant code
(GlobalCache.hpp:14-135, SceneReconstructionEngine_host.cu:417-432, SceneReconstructionEngine.hpp:315-322; Scene(..., useSwapping =
false) at topfu.cpp:67), so the invariance is the parity statement there is."""
import numpy as np
import pytest

from topfusion_b200 import synth

pytestmark = pytest.mark.gpu


def _sweep(n, deg):
    """the camera at the origin turning about its own vertical axis: the S1 scene wanders out of view on one side"""
    scene = synth.scene_s1()
    poses = []
    frames = []
    for i in range(n):
        a = np.radians(deg) * i
        R = synth._rot_y(a)
        p = synth._pose(R, np.zeros(3))
        poses.append(p)
        frames.append(synth.render_depth(scene, p))
    return np.stack(frames), np.stack(poses)


@pytest.fixture(scope="module")
def sweep():
    return _sweep(36, 1.0)


def _blocks(ctx):
    return {k: (v["sdf"].copy(), v["w"].copy()) for k, v in ctx.blocks_by_pos().items()}


def test_streamed_scene_is_identical_to_a_resident_one(gpu, sweep):
    depth, _ = sweep
    plain = gpu.Context(corrected_mode=1)
    streamed = gpu.Context(corrected_mode=1)
    try:
        moved_out = moved_in = 0
        max_store = 0
        for i in range(depth.shape[0]):
            # blocks the coming frame may look at: the enlarged frustum of the last pose (1 degree of motion per frame)
            n_in, left = streamed.stream_in()
            moved_in += n_in
            ok_a = plain.process_frame(depth[i])
            ok_b = streamed.process_frame(depth[i])
            assert ok_a == ok_b, i
            np.testing.assert_array_equal(plain.pose(), streamed.pose(), err_msg=f"frame {i}")
            moved_out += streamed.stream_out()
            max_store = max(max_store, streamed.stream_stats()[1])
        assert moved_out > 500 and max_store > 300, (moved_out, moved_in, max_store)   # the sweep really pushed blocks out
        in_pool, in_store = streamed.stream_stats()
        assert in_pool + in_store == plain.stream_stats()[0]
        assert in_pool < plain.stream_stats()[0]
        # bring everything home and compare voxel for voxel
        n_in, left = streamed.stream_in(all_blocks=True)
        assert left == 0 and streamed.stream_stats() == (plain.stream_stats()[0], 0)
        a, b = _blocks(plain), _blocks(streamed)
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1]), k
    finally:
        plain.close(); streamed.close()


def test_stream_out_frees_pool_slots_and_survives_a_full_pool(gpu, sweep):
    depth, poses = sweep
    small = gpu.Context(corrected_mode=1, num_blocks=4096)
    try:
        for i in range(6):
            small.process_frame(depth[i])
        used0, _ = small.stream_stats()
        # evict whatever is out of view, then look somewhere else: the freed slots are handed out again
        n = small.stream_out()
        used1, stored = small.stream_stats()
        assert stored == n and used1 == used0 - n
        # restoring into a pool that has no room leaves the blocks in the store and says so
        back, left = small.stream_in(all_blocks=True)
        assert back + left == n
        assert small.stream_stats() == (used1 + back, left)
    finally:
        small.close()


def test_reset_drops_the_store_and_save_refuses_while_blocks_are_out(gpu, sweep, tmp_path):
    depth, _ = sweep
    ctx = gpu.Context(corrected_mode=1)
    try:
        for i in range(30):
            ctx.process_frame(depth[i])
        n = ctx.stream_out()
        assert n > 0
        with pytest.raises(Exception):
            ctx.save_scene(str(tmp_path / "x.tfb"))
        ctx.stream_in(all_blocks=True)
        ctx.save_scene(str(tmp_path / "x.tfb"))
        ctx.stream_out()
        ctx.reset()
        assert ctx.stream_stats() == (0, 0)
    finally:
        ctx.close()
