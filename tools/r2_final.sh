#!/bin/bash
# one gpurun call at the end of a round: GPU parity tests, the bench line (both arms), ncu launch list, ncu --set full of the frame's
# kernels (S1 headline frame) and of the scene kernels at the size where bandwidth matters (S3, 2 mm), instrumented-build profiles
set -u
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"
CMD="python tools/profile_frames.py 24 1"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 160 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_icp_all|k_integrate|k_raycast|k_bilateral|k_mark|k_visible_list|k_model_maps|k_pyr_maps|k_points_normals" -s 108 -c 20 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
echo "ncu full rc=$?"
CMD="python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 6 --warmup 3"
$CMD > gpurun_out/plain_micro_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_integrate|k_raycast|k_mark|k_visible_list" -s 12 -c 8 -f -o gpurun_out/prof_micro_$TAG $CMD > gpurun_out/ncu_micro_$TAG.log 2>&1; echo "ncu S3 rc=$?"; tail -1 gpurun_out/plain_micro_$TAG.log
python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 2>&1 | tail -1
python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 --flush-frame 2>&1 | tail -1
python tools/microbench.py --seq S1 --voxel-mm 5 2 --mu-voxels 4 --frames 30 2>&1 | tail -2
for s in "2 0" "4 0" "8 0"; do python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 --shard $s 2>&1 | tail -1; done
TFB_LIB_PATH=topfusion_b200/libtfusion_b200_prof.so python tools/icp_phase_profile.py > gpurun_out/icp_phase_$TAG.log 2>&1; head -22 gpurun_out/icp_phase_$TAG.log
TFB_LIB_PATH=topfusion_b200/libtfusion_b200_prof.so python tools/ray_profile.py 2>&1 | grep -v "^t=\|^warp  " | head -5
python tools/fps_quick.py 90 3 2>&1 | tail -2
python tools/ingest_quick.py 2>&1 | tail -2
python tools/profile_sharded_emulated.py --frames 16 > gpurun_out/shard_emul_$TAG.json 2> gpurun_out/shard_emul_$TAG.err; echo "emul rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "launches", d["gpu_launches"])
print("kernels", {k: round(v["us_per_launch"],1) for k,v in d["kernels"].items()})
print("reference_gpu", {k:v for k,v in d.get("reference_gpu",{}).items() if k.startswith(("debug","as_","this","speed"))})
print("large", d["voxel_updates_large_scene"]["value"], d["voxel_updates_large_scene"]["frac_of_measured_hbm_peak_per_gpu"], d["voxel_updates_large_scene"]["k_integrate_us_slowest_rank"])
print("ingest", d["ingest_from_files"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
