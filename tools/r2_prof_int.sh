#!/bin/bash
# ncu --set full of k_integrate alone on the large scene (S3, 2 mm voxels)
TAG=${1:-int}
CMD="python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 4 --warmup 3"
$CMD > gpurun_out/plain_micro_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_integrate" -s 5 -c 2 -f -o gpurun_out/prof_int_$TAG $CMD > gpurun_out/ncu_int_$TAG.log 2>&1; echo ncu rc=$?; tail -1 gpurun_out/plain_micro_$TAG.log
