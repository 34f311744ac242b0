#!/bin/bash
# ncu --set full of the scene kernels at the size where bandwidth matters (S3, 2 mm voxels, ~49 k visible blocks)
TAG=${1:-s3}
CMD="python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 6 --warmup 3"
$CMD > gpurun_out/plain_micro_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_integrate|k_raycast|k_mark|k_visible_list" -s 12 -c 8 -f -o gpurun_out/prof_micro_$TAG $CMD > gpurun_out/ncu_micro_$TAG.log 2>&1; echo ncu rc=$?; tail -2 gpurun_out/plain_micro_$TAG.log
