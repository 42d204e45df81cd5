#!/bin/bash
# ncu --set full of the dominant kernels (one launch each, after one warm sweep).  Usage: gpu_ncu_full.sh <tag> [regex]
TAG=${1:-prof}
REGEX=${2:-"k_level_sweep|k_level_seg_update|k_spmv|k_predict"}
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1
# skip the first sweep (~180 launches), then take a handful of launches of each kernel from the second
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s 150 -c 12 \
    -o gpurun_out/$TAG -f python tools/profile_step.py --sweeps 3 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/$TAG.ncu-rep
