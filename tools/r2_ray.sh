#!/bin/bash
# one gpurun call for a raycast change: bit-exact ray tests, per-warp lifetimes (instrumented build), per-kernel times on S1 and S3
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scene.py tests/test_gpu_refgpu_fixtures.py tests/test_gpu_sharding.py tests/test_gpu_configs.py -x -q > gpurun_out/ray_tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ray_tests_$TAG.log
TFB_LIB_PATH=topfusion_b200/libtfusion_b200_prof.so python tools/ray_profile.py 2>&1 | grep -v "^t=\|^warp  "
python tools/microbench.py --seq S1 --voxel-mm 5 --mu-voxels 4 --frames 30 2>&1 | tail -1
python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 2>&1 | tail -1
python tools/profile_sharded_emulated.py --frames 16 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items()} if not isinstance(list(d.values())[0],dict) else {k:v for k,v in d.items()})" 2>&1 | cut -c1-1500
