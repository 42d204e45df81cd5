"""Where the wall-clock of MyFMRegressor.fit() goes outside the sweeps (diagnostic): python-side checks,
trainer creation, init_fm, the chain, teardown."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import myfm_b200  # noqa: E402
from myfm_b200 import _myfm  # noqa: E402

wl = bench.Workload(os.environ.get("WORKLOAD", "ml10m"))
with myfm_b200.engine_options(dtype="f32"):  # warm the context
    t = _myfm._TrainerHandle(wl.X[:2000], [], wl.y[:2000], 1, wl.config(2))
    del t
marks = []
orig_create, orig_init, orig_del = _myfm._TrainerHandle.__init__, _myfm._TrainerHandle.init_fm, _myfm._TrainerHandle.__del__


def timed(name, fn):
    def wrapper(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        marks.append((name, time.perf_counter() - t0))
        return out
    return wrapper


_myfm._TrainerHandle.__init__ = timed("trainer create", orig_create)
_myfm._TrainerHandle.init_fm = timed("init_fm", orig_init)
_myfm._TrainerHandle.__del__ = timed("trainer destroy", orig_del)
stamps = []
t0 = time.perf_counter()
with myfm_b200.engine_options(dtype="f32"):
    fm = myfm_b200.MyFMRegressor(rank=wl.rank, random_seed=42).fit(
        wl.X, wl.y, n_iter=13, n_kept_samples=1, group_shapes=wl.group_shapes,
        callback=lambda i, f, h, hist: (stamps.append(time.perf_counter()), (False, None))[1])
t1 = time.perf_counter()
print(f"fit total {t1 - t0:.3f} s; first callback at {stamps[0] - t0:.3f} s; chain {stamps[-1] - stamps[0]:.3f} s; "
      f"after the last callback {t1 - stamps[-1]:.3f} s")
for name, dt in marks:
    print(f"  {name}: {dt:.3f} s")
