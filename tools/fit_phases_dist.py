"""Where the wall-clock of a row-sharded MyFMRegressor.fit() goes before the first sweep (diagnostic).

    MYFM_TRACE_SETUP=1 torchrun --nproc-per-node 2 tools/fit_phases_dist.py      (rank 0 prints)
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import myfm_b200  # noqa: E402
from myfm_b200 import _myfm, distributed as mdist  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
wl = bench.Workload("ml10m")
t0 = time.perf_counter()
X, y, ctx = mdist.shard(wl.X, wl.y)
t_shard = time.perf_counter() - t0
marks = []


def timed(name, fn):
    def wrapper(*a, **k):
        t = time.perf_counter()
        out = fn(*a, **k)
        marks.append((name, time.perf_counter() - t))
        return out
    return wrapper


_myfm._TrainerHandle.__init__ = timed("trainer create (C ABI)", _myfm._TrainerHandle.__init__)
_myfm._TrainerHandle.init_fm = timed("init_fm", _myfm._TrainerHandle.init_fm)
mdist._fresh_unique_id = timed("fresh ncclUniqueId (broadcast)", mdist._fresh_unique_id)
for attempt in range(2):
    marks.clear()
    stamps = []
    dist.barrier()
    t0 = time.perf_counter()
    with ctx.options(dtype="f32"):
        t_opt = time.perf_counter() - t0
        myfm_b200.MyFMRegressor(rank=wl.rank, random_seed=42).fit(
            X, y, n_iter=8, n_kept_samples=1, group_shapes=wl.group_shapes,
            callback=lambda i, f, h, hist: (stamps.append(time.perf_counter()), (False, None))[1])
    t1 = time.perf_counter()
    if rank == 0:
        print(f"fit #{attempt}: shard() {t_shard:.3f} s (outside); options() {t_opt:.3f} s; first callback at "
              f"{stamps[0] - t0:.3f} s; chain {stamps[-1] - stamps[0]:.3f} s; after {t1 - stamps[-1]:.3f} s", flush=True)
        for name, dt in marks:
            print(f"    {name}: {dt:.3f} s", flush=True)
dist.destroy_process_group()
