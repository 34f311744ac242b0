#!/bin/bash
# short evidence refresh on the final tree: GPU tests, bench line (both arms), ncu launch list
set -u
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
CMD="python tools/profile_frames.py 24 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 160 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "ncu list rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "warm", d["warm_l2_value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic_source"])
print("kernels", {k: round(v["us_per_launch"],1) for k,v in d["kernels"].items()})
print("large", d["voxel_updates_large_scene"]["value"], d["voxel_updates_large_scene"]["frac_of_measured_hbm_peak_per_gpu"])
print("ingest", d["ingest_from_files"]["value"], "refgpu", d["reference_gpu"]["speedup_vs_debug_work_removed"])
PY
