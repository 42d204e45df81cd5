#!/bin/bash
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-300
