#!/bin/bash
# one `gpurun --gpus 8` call: bit-for-bit parity of the sharded scene against a single context with real processes, then the bench line
N=${1:-8}; TAG=${2:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500"
timeout 200 $TR tools/check_multigpu.py > gpurun_out/mg_check_${N}_$TAG.log 2>&1; echo "check rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/mg_check_${N}_$TAG.log | tail -3
timeout 600 $TR bench.py --gpus $N --steps 60 --warmup 5 > gpurun_out/bench_g${N}_$TAG.json 2> gpurun_out/bench_g${N}_$TAG.err; echo "bench rc=$?"
timeout 300 $TR bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ref_g${N}_$TAG.json 2>> gpurun_out/bench_g${N}_$TAG.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_g${N}_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k in ("voxel_updates_large_scene","voxel_updates_large_scene_1280x720"):
    print(k, d[k]["value"], d[k]["k_integrate_us_slowest_rank"], d[k]["visible_blocks_per_frame_all_ranks"])
r=json.load(open("gpurun_out/bench_ref_g${N}_$TAG.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
