#!/bin/bash
# parity tests + diag timings + launch list
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-400
python tools/diag.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --sweeps 3 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
