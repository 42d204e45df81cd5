"""Timing diagnostics on the bench workload: sweeps with/without the per-family event timers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, myfm_b200
from myfm_b200._myfm import ConfigBuilder, _TrainerHandle

wl = sys.argv[1] if len(sys.argv) > 1 else "ml10m"
X, y, gs, rank = bench.make_workload(wl)
cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(gs)), gs)).set_n_iter(100).set_n_kept_samples(1).build())
with myfm_b200.engine_options(dtype=os.environ.get("DTYPE", "f32")):
    t = _TrainerHandle(X, [], y, 42, cfg); t.init_fm(rank, 0.1)
t.step(3); t.sync()
for rep in range(3):
    ms = t.timed_steps(10)
    print(f"no-profiling: {ms/10:.3f} ms/sweep", flush=True)
t.set_profiling(True)
ms = t.timed_steps(10)
print(f"profiling: {ms/10:.3f} ms/sweep; families:", [t.kernel_ms(f)[0] / 10 for f in range(3)], flush=True)
