#!/bin/bash
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1
python tools/diag.py 2>&1 | tail -5
MYFM_HOST_RNG=1 python tools/diag.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --sweeps 3 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
