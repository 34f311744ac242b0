#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): the scene sharded over the ranks with CUDA-IPC peer pointers and NCCL barriers
must reproduce, bit for bit, a single context on rank 0's GPU that holds the whole scene (poses, raycast, voxels).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 tools/check_multigpu.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from topfusion_b200 import capi, multigpu, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    n = int(os.environ.get("TFB_CHECK_FRAMES", "12"))
    depth, _, _ = synth.sequence("S1", n)
    eng = multigpu.CudaEngine(rank, world, local, dist, plumbing=os.environ.get("TFB_PLUMBING", "p2p"), corrected_mode=1)
    single = capi.Context(corrected_mode=1) if rank == 0 else None
    bad = 0
    with torch.cuda.stream(eng.stream):
        buf = torch.empty(depth.shape[1:], dtype=torch.int16, device=torch.device("cuda", local))
        st = multigpu.ShardedTopFu(eng, dist, rank, world, buf)
        for i in range(n):
            src = torch.from_numpy(depth[i].view(np.int16)).pin_memory() if rank == 0 else None
            ok = st.process_frame(src)
            eng.ctx.sync()                                   # collective: finishes the deferred tail on every rank
            upd = st.total(eng.voxel_updates())
            if rank == 0:
                ok1 = single.process_frame(depth[i])
                same_pose = np.array_equal(eng.pose().view(np.uint32), single.pose().view(np.uint32))
                same_ray = i == 0 or np.array_equal(eng.ctx.raycast_result().view(np.uint32), single.raycast_result().view(np.uint32))
                same_upd = upd == single.voxel_updates()
                print(f"frame {i}: ok {ok}/{ok1} pose {'==' if same_pose else '!='} raycast {'==' if same_ray else '!='} "
                      f"updates {'==' if same_upd else '!='} ({single.voxel_updates()})", flush=True)
                bad += int(not (ok == ok1 and same_pose and same_ray and same_upd))
        # every rank's replica of the pose must be the same bits
        p = torch.from_numpy(eng.pose().view(np.int32).astype(np.int64)).cuda()
        lo, hi = p.clone(), p.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        bad += int(not torch.equal(lo, hi))
        owned = st.total(len(eng.ctx.blocks_by_pos()))
    if rank == 0:
        print(f"blocks owned over all ranks: {owned:.0f}; single context: {len(single.blocks_by_pos())}; mismatches: {bad}")
        single.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
