#!/bin/bash
# Quick GPU visit: parity tests, then sweep timing with per-family CUDA-event times.  Usage: gpu_quick.sh [pytest -k expr]
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests -m gpu -q ${1:+-k "$1"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_step.py --sweeps 10 --families 2>&1 | tail -5
