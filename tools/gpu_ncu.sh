#!/bin/bash
# ncu launch list + full-set capture of the hot kernels.  Usage: gpu_ncu.sh <tag> [kernel regex] [skip] [count]
TAG=${1:-prof}
REGEX=${2:-"k_field_stream|k_field_stats|k_predict"}
SKIP=${3:-120}
COUNT=${4:-6}
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py --sweeps 1 > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT \
    -o gpurun_out/full_$TAG -f python tools/profile_step.py --sweeps 1 > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out/
