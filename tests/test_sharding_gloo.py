"""Multi-rank host logic of the sharded scene on the CPU: world size 2, gloo.

topfusion_b200.multigpu.ShardedTopFu is driven with an engine built on the ORACLE (shard_rank / shard_count are part
of its parameters too): the frame is broadcast from rank 0, every rank allocates the replicated index and integrates
only the blocks it owns, and the union of the shards must equal the single-rank scene voxel for voxel.  The partition
function of the host side (multigpu.owner_rank) must agree with the one compiled into the oracle and the kernels."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """begin/raycast/end over the oracle's stage API with injected poses (allocation + integration: the sharded stages)"""

    def __init__(self, rank, world, poses_w2c, cols, rows):
        from oracle import tfo
        self.L = tfo.Lib("port")
        self.o = tfo.Oracle(lib=self.L, shard_rank=rank, shard_count=world, cols=cols, rows=rows,
                            fx=504.261 * cols / 640, fy=503.905 * rows / 480, cx=352.457 * cols / 640, cy=272.202 * rows / 480)
        self.poses = poses_w2c
        self.i = 0
        self.updates = 0

    def begin(self, frame):
        depth = frame.numpy().view(np.uint16)
        dists = self.L.compute_dists(depth)
        self.o.allocate(self.poses[self.i], dists)
        self.o.integrate(self.poses[self.i], dists)
        self.updates = self.o.voxel_updates()
        self.i += 1

    def raycast(self):
        pass

    def end(self):
        return True

    def voxel_updates(self):
        return self.updates


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from topfusion_b200 import multigpu, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cols, rows, n = 160, 120, 4
    depth, poses, _ = synth.sequence("S1", n, cols, rows)
    w2c = [np.linalg.inv(p).astype(np.float32) for p in poses]
    eng = OracleEngine(rank, world, w2c, cols, rows)
    buf = torch.zeros((rows, cols), dtype=torch.int16)
    st = multigpu.ShardedTopFu(eng, dist, rank, world, buf)
    totals = []
    for i in range(n):
        src = torch.from_numpy(depth[i].view(np.int16)) if rank == 0 else None   # only rank 0 holds the frames
        assert st.process_frame(src)
        totals.append(st.total(eng.voxel_updates(), "sum"))
    t = eng.o.table()
    blocks = eng.o.blocks_by_pos()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
             alloc=np.array(sorted({tuple(int(v) for v in e) for e in t[t["ptr"] >= -1]["pos"]}), dtype=np.int32).reshape(-1, 3),
             pos=np.array(sorted(blocks), dtype=np.int32).reshape(-1, 3),
             vox=np.stack([np.stack([blocks[k]["sdf"].astype(np.int32), blocks[k]["w"].astype(np.int32)]) for k in sorted(blocks)]),
             totals=np.array(totals))
    dist.barrier()
    dist.destroy_process_group()


def test_union_of_shards_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from topfusion_b200 import multigpu, synth
    from oracle import tfo
    world = 2
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

    cols, rows, n = 160, 120, 4
    depth, poses, _ = synth.sequence("S1", n, cols, rows)
    L = tfo.Lib("port")
    o = tfo.Oracle(lib=L, cols=cols, rows=rows, fx=504.261 * cols / 640, fy=503.905 * rows / 480, cx=352.457 * cols / 640,
                   cy=272.202 * rows / 480)
    single_totals = []
    for i in range(n):
        w2c = np.linalg.inv(poses[i]).astype(np.float32)
        d = L.compute_dists(depth[i])
        o.allocate(w2c, d)
        o.integrate(w2c, d)
        single_totals.append(o.voxel_updates())
    ref = o.blocks_by_pos()
    ref_alloc = tfo.allocated_set(o.table())
    o.close()

    shards = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    union = {}
    for r, z in enumerate(shards):
        assert {tuple(int(v) for v in p) for p in z["alloc"]} == ref_alloc, "the index must be a replica on every rank"
        for p, v in zip(z["pos"], z["vox"]):
            key = tuple(int(x) for x in p)
            assert key not in union, "a block has two owners"
            assert multigpu.owner_rank(*key, world) == r, "host-side partition function disagrees with the oracle"
            union[key] = v
        assert list(z["totals"]) == [float(t) for t in single_totals], "aggregated voxel-updates must equal one rank's"
    assert set(union) == set(ref)
    for k, v in ref.items():
        assert np.array_equal(union[k][0], v["sdf"].astype(np.int32)) and np.array_equal(union[k][1], v["w"].astype(np.int32)), k
    sizes = [len(z["pos"]) for z in shards]
    assert min(sizes) > 0.35 * len(ref), sizes


def test_row_partition_covers_the_image():
    from topfusion_b200 import multigpu
    for rows in (480, 720, 120):
        for n in (1, 2, 3, 4, 8):
            got = sorted(r for k in range(n) for r in multigpu.rows_of_rank(rows, k, n))
            assert got == list(range(rows))
