#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): where the sharded frame spends its time.  Per-kernel CUDA-event times (the
library's own ktiming) of the collective pipelined frame on every rank; a barrier's time includes the wait for the slowest
rank, which is the point.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29500 tools/shard_kernel_times.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from topfusion_b200 import multigpu, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    n, warm = int(os.environ.get("TFB_FRAMES", "40")), 8
    depth, _, _ = synth.sequence("S1", n + warm)
    eng = multigpu.CudaEngine(rank, world, local, dist, plumbing="p2p", corrected_mode=1)
    with torch.cuda.stream(eng.stream):
        buf = torch.empty(depth.shape[1:], dtype=torch.int16, device=torch.device("cuda", local))
        st = multigpu.ShardedTopFu(eng, dist, rank, world, buf)
        for i in range(n + warm):
            if i == warm:
                eng.ctx.sync()
                eng.ctx.ktiming(True)
            src = torch.from_numpy(depth[i].view(np.int16)).pin_memory() if rank == 0 else None
            st.process_frame(src)
        eng.ctx.sync()
        times = eng.ctx.kernel_times()
    mine = {k: round(1000.0 * ms / n, 2) for k, (ms, cnt) in times.items()}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        out = {"frames": n, "ranks": world, "us_per_frame_by_rank": allr,
               "sum_us_by_rank": [round(sum(r.values()), 1) for r in allr]}
        print(json.dumps(out))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
