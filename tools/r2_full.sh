#!/bin/bash
# one gpurun call: the whole GPU test suite, the bench line (both arms)
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
print("kernels", {k: round(v["us_per_launch"],1) for k,v in d["kernels"].items()})
print("reference_gpu", d.get("reference_gpu"))
print("large", d["voxel_updates_large_scene"])
print("ingest", d["ingest_from_files"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
