"""Minimal driver for ncu: builds the bench workload and runs a few sweeps (no e2e / CPU legs).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py --sweeps 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import myfm_b200  # noqa: E402
from myfm_b200._myfm import _TrainerHandle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="ml10m")
ap.add_argument("--sweeps", type=int, default=2)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--families", action="store_true", help="per-family CUDA-event times")
ap.add_argument("--shard", default="", help="R/N: only the rows rank R of N would hold (column partition), on one GPU")
args = ap.parse_args()

wl = bench.Workload(args.workload)
if args.shard:  # what one rank of a row-sharded run computes, without its peers (kernel times under ncu)
    from myfm_b200.distributed import column_partition
    r, n = (int(v) for v in args.shard.split("/"))
    wl.X.sort_indices()
    rows = (column_partition(wl.X, n) == r).nonzero()[0]
    wl.X, wl.y = wl.X[rows], wl.y[rows]
    print(f"shard {r}/{n}: {rows.size} rows")
with myfm_b200.engine_options(dtype=args.dtype):
    t = _TrainerHandle(wl.X, wl.blocks(), wl.y_engine, bench.CHAIN_SEED, wl.config(args.sweeps))
    t.init_fm(wl.rank, 0.1)
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
