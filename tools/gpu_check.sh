#!/bin/bash
# one gpurun call: GPU parity tests, the bench line, then the ncu launch list and a full capture of the top kernels
set -u
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_icp_all|k_integrate$|k_raycast|k_bilateral|k_mark|k_visible_list|k_model_maps|k_pyr_maps" -s 16 -c 16 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
# the sharded kernels on ONE GPU (two emulated ranks: every peer access is local) — the only legitimate way to put ncu on them
CMD2="python tools/profile_sharded_emulated.py --frames 16"
$CMD2 > gpurun_out/shard_emul_$TAG.json 2> gpurun_out/shard_emul_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:"k_raycast_sharded|k_gather_foreign" -s 24 -c 4 -f -o gpurun_out/prof_shard_emul_$TAG $CMD2 > gpurun_out/ncu_shard_emul_$TAG.log 2>&1
echo "ncu sharded (emulated) rc=$?"
