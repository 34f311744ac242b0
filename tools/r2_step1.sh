#!/bin/bash
python -m pytest tests/test_gpu_refgpu_fixtures.py tests/test_gpu_scene.py -q > gpurun_out/r2_fixt2.log 2>&1; echo rc=$?; tail -15 gpurun_out/r2_fixt2.log
python -c "
import json; d=json.load(open('gpurun_out/refgpu_fixture_parity.json')); print({k:v for k,v in d.items() if k.startswith(('voxels','identity','blocks','model'))})"
python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 --out gpurun_out/r2_micro_s3_dev.json
python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 --ieee 1
python tools/microbench.py --seq S1 --voxel-mm 5 2 --mu-voxels 4 --frames 30
