"""Minimal driver for ncu: builds the bench workload and runs a few sweeps (no e2e / CPU legs).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --sweeps 2
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import myfm_b200  # noqa: E402
from myfm_b200._myfm import ConfigBuilder, _TrainerHandle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ml10m")
ap.add_argument("--sweeps", type=int, default=2)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--families", action="store_true", help="per-family CUDA-event times")
args = ap.parse_args()

X, y, group_shapes, rank = bench.make_workload(args.workload)
cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(group_shapes)), group_shapes))
       .set_n_iter(args.sweeps).set_n_kept_samples(1).build())
with myfm_b200.engine_options(dtype=args.dtype):
    t = _TrainerHandle(X, [], y, bench.CHAIN_SEED, cfg)
    t.init_fm(rank, 0.1)
t.timed_steps(3)  # warm-up
if args.families:
    t.set_profiling(True)
ms = t.timed_steps(args.sweeps)
if args.families:
    for fam, name in enumerate(("column_sweeps", "q_init", "e_refresh", "  stream level", "  gather level")):
        fms, n = t.kernel_ms(fam)
        print(f"  {name}: {fms / args.sweeps:.3f} ms/sweep in {n // args.sweeps} launch groups")
    t.set_profiling(False)
    t.timed_steps(2)  # captures the CUDA graphs
    ms_graph = t.timed_steps(args.sweeps)
    print(f"with CUDA-graph replay: {ms_graph / args.sweeps:.3f} ms/sweep")
print(f"sweep path = {t.sweep_path()}")
print(f"{args.sweeps} sweeps: {ms / args.sweeps:.2f} ms/sweep, launches={t.launch_count()}")
