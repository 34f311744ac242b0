#!/bin/bash
# one gpurun call for an integration change: voxel-exact tests, then k_integrate on the large scene and the headline frame
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scene.py tests/test_gpu_refgpu_fixtures.py tests/test_gpu_sharding.py tests/test_gpu_configs.py tests/test_gpu_pipeline.py -x -q > gpurun_out/int_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/int_tests_$TAG.log
python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 2>&1 | tail -1
python tools/microbench.py --seq S1 --voxel-mm 5 2 --mu-voxels 4 --frames 30 2>&1 | tail -2
