#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms + other workloads), ncu launch list, ncu --set full of the hot kernels.
# Usage: tools/gpu_round.sh [tag]     outputs under gpurun_out/
TAG=${1:-round}
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --dtype f64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_f64.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --workload ml1m-ext --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_ml1mext.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --workload ml100k --steps 50 --warmup 5 > gpurun_out/bench_${TAG}_ml100k.json 2>> gpurun_out/bench.err
for t in classification ordered; do for r in mt19937 philox; do
  timeout 300 python bench.py --workload ml100k --task $t --rng $r --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_ml100k_${t}_$r.json 2>> gpurun_out/bench.err
done; done
for f in gpurun_out/bench_${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(l["value"],2), "it/s", "e2e", round(l["e2e"]["value"],2), "cpu", (l.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
# launch list of the bench command itself (a number printed under ncu is never a bench value)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
# full-set capture of the hot kernels; the reports stay on the box, their raw pages come back as CSV
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_field_stream|k_field_stats" -s 40 -c 4 \
    -o /tmp/full_$TAG -f python tools/profile_step.py --sweeps 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_predict_tile|k_mt_farm|k_group_hyper|k_reduce_e_both" -s 4 -c 5 \
    -o /tmp/full2_$TAG -f python tools/profile_step.py --sweeps 3 > gpurun_out/ncu_full2.log 2>&1
ncu -i /tmp/full2_$TAG.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_other_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -5
