CMD="python tools/microbench.py --voxel-mm 2 --mu-voxels 8 --frames 8 --warmup 3"
$CMD > gpurun_out/plain_micro.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_integrate|k_raycast" -s 12 -c 4 -f -o gpurun_out/prof_micro $CMD > gpurun_out/ncu_micro.log 2>&1; echo rc=$?; tail -3 gpurun_out/plain_micro.log
