"""Observed parity of the f32 / f64 engines with the oracle, as max ELEMENTWISE relative errors (what the
tolerances of tests/test_gpu_parity.py leave room for): per array and sweep, max |engine - oracle| / |oracle| over
the entries with |oracle| above 1e-3 of the array's largest magnitude, and max |diff| / max |oracle| over all.
Free-running chains on the same seed, and (f32) teacher-forced sweeps restarted from the oracle's state.

    python tools/parity_report.py [--rows 80000] [--sweeps 20]      -> markdown on stdout (profiles/r02_parity_report.md)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import myfm_b200  # noqa: E402
from helpers import movielens_like  # noqa: E402
from myfm_b200._myfm import FMHyperParameters  # noqa: E402
from oracle import binding as oracle  # noqa: E402
from test_gpu_parity import make_pair  # noqa: E402


def errors(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    scale = float(np.max(np.abs(b))) if b.size else 1.0
    big = np.abs(b) > 1e-3 * scale
    rel_big = float(np.max(np.abs(a - b)[big] / np.abs(b)[big])) if big.any() else 0.0
    return rel_big, float(np.max(np.abs(a - b)) / max(scale, 1e-300))


def state(trainer, chain):
    w0, w, V, _ = trainer.get_fm()
    ow0, ow, oV = chain.fm()
    h, oh = trainer.get_hyper(), chain.hyper()
    return {"w0": ([w0], [ow0]), "w": (w, ow), "V": (V, oV), "alpha": ([h.alpha], [oh["alpha"]]),
            "lambda_V": (h.lambda_V, oh["lambda_V"]), "mu_V": (h.mu_V, oh["mu_V"]), "e": (trainer.get_e(), chain.e())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=80000)
    ap.add_argument("--sweeps", type=int, default=20)
    args = ap.parse_args()
    oracle.build()
    X, y, gs = movielens_like(args.rows, 943, 1682, 8, seed=0)
    y = np.clip(np.round(y), 1, 5)
    print(f"# Observed parity with the oracle (ml100k-shaped, {args.rows} rows, rank 8, {args.sweeps} sweeps, same seed)\n")
    print("Per array: worst over the sweeps of  max |engine - oracle| / |oracle|  over the entries above 1e-3 of the "
          "array's largest magnitude  /  max |engine - oracle| / max |oracle|  over all entries.\n")
    for dtype, forced in (("f64", False), ("f32", False), ("f32", True)):
        trainer, chain = make_pair(myfm_b200, oracle, X, y, 8, dtype, group_shapes=gs, n_iter=args.sweeps)
        worst = {}
        for it in range(args.sweeps):
            if forced and it > 0:  # restart the engine from the oracle's state: isolates one sweep's error
                ow0, ow, oV = chain.fm()
                oh = chain.hyper()
                trainer.set_state(ow0, ow, oV, FMHyperParameters(oh["alpha"], oh["mu_w"], oh["lambda_w"], oh["mu_V"],
                                                                 oh["lambda_V"]), chain.e())
            trainer.step(1)
            chain.step()
            for name, (a, b) in state(trainer, chain).items():
                r = errors(a, b)
                worst[name] = tuple(max(x, y_) for x, y_ in zip(worst.get(name, (0.0, 0.0)), r))
        mode = "teacher-forced (every sweep starts from the oracle's state)" if forced else "free-running"
        print(f"## {dtype}, {mode}\n\n| array | max rel. error, entries above 1e-3 of the scale | max abs. error / scale |\n|---|---|---|")
        for name, (r1, r2) in worst.items():
            print(f"| {name} | {r1:.2e} | {r2:.2e} |")
        print()


if __name__ == "__main__":
    main()
