#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the hot kernels.
# Usage: tools/gpu_round.sh [tag]     outputs under gpurun_out/
TAG=${1:-round}
set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
# launch list of the bench command itself (a number printed under ncu is never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
# full-set capture of the hot kernels; the report stays on the box, its raw page comes back as CSV
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_field_stream|k_field_stats|k_predict_tile" -s 40 -c 5 \
    -o /tmp/full_$TAG -f python tools/profile_step.py --sweeps 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/full_$TAG.ncu-rep
