#!/bin/bash
# one `gpurun --gpus N` call: bit-for-bit parity of the sharded scene against a single context (real processes, CUDA IPC, the
# library's own cross-GPU flags), the bench line at N GPUs, and where the sharded frame spends its time.  Never under ncu.
#   gpurun --gpus 2 --timeout 500 -- 'bash tools/multigpu_check.sh 2 tag'
set -u
N=${1:-2}
TAG=${2:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 200 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_multiprocess.py -m gpu -x -q 2>&1 | tail -2
timeout 150 $TR tools/check_multigpu.py > gpurun_out/mg_check_${N}_$TAG.log 2>&1; echo "check rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/mg_check_${N}_$TAG.log | tail -2
timeout 500 $TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/bench_g${N}_$TAG.json 2> gpurun_out/bench_g${N}_$TAG.err; echo "bench rc=$?"
timeout 120 $TR tools/shard_kernel_times.py > gpurun_out/ktimes_g${N}_$TAG.json 2> gpurun_out/ktimes_g${N}_$TAG.err; echo "kernel times rc=$?"
