#!/bin/bash
# k_integrate variants built into topfusion_b200/_build/lib_<name>.so: the large scene and the headline frame for each
echo "== default"; python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 2>&1 | tail -1
python tools/microbench.py --seq S1 --voxel-mm 5 2 --mu-voxels 4 --frames 30 2>&1 | tail -2
for f in topfusion_b200/_build/lib_*.so; do
  echo "== $f"
  TFB_LIB_PATH=$f python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30 2>&1 | tail -1
  TFB_LIB_PATH=$f python tools/microbench.py --seq S1 --voxel-mm 5 2 --mu-voxels 4 --frames 30 2>&1 | tail -2
done
