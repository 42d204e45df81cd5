"""Where the host time of one fit() iteration goes (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import myfm_b200  # noqa: E402
from myfm_b200._myfm import ConfigBuilder, _TrainerHandle, _LiveFM  # noqa: E402

X, y, gs, rank = bench.make_workload("ml10m")
cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(gs)), gs)).set_n_iter(30)
       .set_n_kept_samples(1).build())
t0 = time.perf_counter()
with myfm_b200.engine_options(dtype=os.environ.get("DTYPE", "f32")):
    t = _TrainerHandle(X, [], y, 42, cfg)
    t1 = time.perf_counter()
    t.init_fm(rank, 0.1)
t2 = time.perf_counter()
print(f"trainer create {t1 - t0:.3f} s, init_fm {t2 - t1:.3f} s")
acc = np.zeros(4)
for it in range(30):
    a = time.perf_counter()
    t.step(1)
    b = time.perf_counter()
    t.sync()
    c = time.perf_counter()
    h = t.get_hyper()
    d = time.perf_counter()
    w0 = _LiveFM(t).w0
    e = time.perf_counter()
    if it >= 10:
        acc += [b - a, c - b, d - c, e - d]
acc /= 20
print("per iteration (ms): step() enqueue %.3f, wait for sweep %.3f, get_hyper %.3f, w0 %.3f, total %.3f" %
      (*(acc * 1e3), acc.sum() * 1e3))
