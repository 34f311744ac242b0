#!/bin/bash
M="python tools/microbench.py --seq S3 --voxel-mm 2 --mu-voxels 8 --frames 30"
echo "== skip, dirty flush"; $M 2>&1 | tail -1
echo "== skip, clean flush"; TFB_FLUSH_READBACK=1 $M 2>&1 | tail -1
echo "== noskip, clean flush"; TFB_FLUSH_READBACK=1 TFB_LIB_PATH=topfusion_b200/_build/lib_noskip.so $M 2>&1 | tail -1
echo "== skip, flush per frame"; $M --flush-frame 2>&1 | tail -1
echo "== skip, clean flush per frame"; TFB_FLUSH_READBACK=1 $M --flush-frame 2>&1 | tail -1
echo "== shard 8"; TFB_FLUSH_READBACK=1 $M --shard 8 0 2>&1 | tail -1; $M --shard 8 0 2>&1 | tail -1
