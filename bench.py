#!/usr/bin/env python
"""Headline benchmark: Gibbs iterations/s of the Bayesian-FM sweep at rank 32 on the
MovieLens-10M-shaped workload (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  A "step" is one full update_all sweep (alpha, w0, hypers, w,
K factor columns of V, e refresh) over the whole training set.

value     whole-job iterations/s, inputs resident in HBM, timed with CUDA events on the engine's
          stream around exactly K sweeps (barrier + synchronize on both sides, max over ranks).
e2e       the same metric through the public API (MyFMRegressor.fit with host scipy/numpy
          buffers): wall clock between the per-iteration callbacks, which includes every
          host->device copy of the sweep's variates (pinned) and the device->host read of the
          sweep's hyper-parameters.
roofline  column-sweep kernels (the dominant family): algorithmic bytes / CUDA-event time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (rows, users, movies, rank)
    "ml10m": (10_000_054, 69_878, 10_677, 32),
    "ml1m": (900_188, 6_040, 3_706, 32),
    "ml100k": (80_000, 943, 1_682, 8),
}
DATA_SEED = 2
CHAIN_SEED = 42


def make_workload(name: str):
    from helpers import movielens_like

    rows, users, movies, rank = WORKLOADS[name]
    # popularity exponents matched to ML-10M's public max/mean ratings per user (7359/143) and
    # per movie (34864/937)
    X, y, group_shapes = movielens_like(rows, users, movies, 8, seed=DATA_SEED, zipf=(0.4, 0.45))
    y = np.clip(np.round(y * 2) / 2, 0.5, 5.0)  # half-star grid like ML-10M
    return X, y, group_shapes, rank


def algorithmic_bytes(nnz: int, n_rows: int, rank: int, real_bytes: int = 4):
    """bytes one sweep must move (i32 indices, `real_bytes` values)."""
    b = real_bytes
    v_sweep = (4 + b) * nnz + 4 * b * nnz          # per factor: CSC read + gather/scatter of q, e
    q_init = (4 + b) * nnz + 4 * n_rows + b * n_rows
    w_sweep = (4 + b) * nnz + 2 * b * nnz
    e_refresh = (4 + b) * nnz + 4 * n_rows + 2 * b * n_rows
    misc = 3 * b * n_rows
    return dict(sweeps=rank * v_sweep + w_sweep, total=rank * (v_sweep + q_init) + w_sweep + e_refresh + misc)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.samples, self._stop_evt = device, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons}


def oracle_chain(X, y, group_shapes, rank, dtype="f32"):
    from oracle import binding as oracle

    oracle.build()
    return oracle.OracleChain(X, y, rank, dtype=dtype, seed=CHAIN_SEED, group_shapes=group_shapes, n_iter=10 ** 6)


def cpu_baseline(X, y, group_shapes, rank, budget_s: float = 20.0):
    """The reference's CPU path (oracle port, f32, 1 thread) on the same workload; bounded."""
    chain = oracle_chain(X, y, group_shapes, rank)
    t1 = chain.timed_steps(1)  # warm-up, also sizes the sample
    n = int(max(1, min(20, budget_s // max(t1, 1e-3))))
    t = chain.timed_steps(n)
    return {"value": n / t, "unit": "it/s", "cores": 1, "kind": "port",
            "sample": f"{n} full update_all sweeps (after 1 warm-up) of the same workload, oracle restatement of the "
                      f"reference sampler (Real=float, g++ -O3, single thread as the reference is; host has "
                      f"{os.cpu_count()} cores)"}


def ncu_traffic(rank: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the two column-sweep kernels from the newest
    committed `ncu --set full` capture (profiles/*_full_raw.csv), per sweep: (rank + 1) launches each."""
    import csv
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_full_raw.csv")))
    if not files:
        return None, None
    rows = list(csv.reader(open(files[-1])))
    hdr, units = rows[0], rows[1]
    try:
        i_name, i_rd, i_wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = {}
    for r in rows[2:]:
        for key in ("k_field_stream", "k_field_stats"):
            if key in r[i_name]:
                b = float(r[i_rd]) * scale.get(units[i_rd], 1.0) + float(r[i_wr]) * scale.get(units[i_wr], 1.0)
                per_kernel.setdefault(key, []).append(b)
    if len(per_kernel) < 2:
        return None, None
    per_vector = sum(sum(v) / len(v) for v in per_kernel.values())
    return per_vector * (rank + 1), os.path.relpath(files[-1], ROOT)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle port; the reference cannot be
    built without Eigen) on the host cores it can use (one: the sampler is single-threaded)."""
    rank_id = int(os.environ.get("RANK", "0"))
    if rank_id != 0:
        return
    X, y, group_shapes, rank = make_workload(args.workload)
    chain = oracle_chain(X, y, group_shapes, rank)
    # bounded: the full workload every step, but only as many steps as fit in ~3 minutes
    t1 = chain.timed_steps(1)
    budget = 180.0
    warm = max(0, min(args.warmup - 1, int(0.2 * budget // max(t1, 1e-6))))
    if warm:
        chain.timed_steps(warm)
    steps = int(max(1, min(args.steps, (budget - (1 + warm) * t1) // max(t1, 1e-6))))
    t = chain.timed_steps(steps)
    value = steps / t
    line = {
        "impl": "reference", "metric": "gibbs_iterations_per_sec", "value": value, "unit": "it/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / steps,
        "steps_run": steps, "warmup_run": 1 + warm,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, X.shape[0], X.nnz, rank, args.gpus),
        "nnz_rank_per_sec": value * X.nnz * rank,
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} full update_all sweeps (of {args.steps} asked; bounded to ~3 min) of the same "
                                   f"workload after {1 + warm} warm-up, oracle restatement of the reference sampler "
                                   f"(Real=float, g++ -O3), 1 thread: the reference sampler is single-threaded "
                                   f"(host has {os.cpu_count()} cores)"},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(name, n_rows, nnz, rank, n_gpus):
    rows, users, movies, _ = WORKLOADS[name]
    return {"workload": f"{name}-shaped synthetic: {rows} rows x ({users} users + {movies} movies) one-hot, "
                        f"nnz={nnz}, rank {rank}, regression, group_shapes=[users, movies]",
            "rank": rank, "rows": int(n_rows), "nnz": int(nnz),
            "rng": "mt19937 (reference stream, same seed; generated on the device)",
            "l2": "working set (CSR+CSC+{e,q},y ~ 0.5 GB) exceeds the 126 MB L2; no explicit flush",
            "parallelism": "single GPU" if n_gpus == 1 else
            f"rows sharded over {n_gpus} GPUs (contiguous ranges), one NCCL all-reduce of the column statistics "
            f"per dependency level, model replicated"}


def run_ours(args):
    import myfm_b200
    from myfm_b200._myfm import ConfigBuilder, _TrainerHandle

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank_id = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    X, y, group_shapes, rank = make_workload(args.workload)
    n_rows_global, nnz_global = X.shape[0], X.nnz
    dtype = args.dtype
    cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(group_shapes)), group_shapes))
           .set_n_iter(args.steps + args.warmup).set_n_kept_samples(1).build())
    if dist is not None:  # every rank generated the same data; keep this rank's contiguous row range
        from myfm_b200 import distributed as mdist

        X, y, ctx = mdist.shard(X, y)
        options = lambda: ctx.options(dtype=dtype, device=local)  # noqa: E731
    else:
        options = lambda: myfm_b200.engine_options(dtype=dtype, device=local)  # noqa: E731

    # ---- device-resident leg ---------------------------------------------------------------
    with options():
        trainer = _TrainerHandle(X, [], y, CHAIN_SEED, cfg)
        trainer.init_fm(rank, 0.1)
    trainer.step(args.warmup)
    trainer.sync()
    launches0 = trainer.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    if dist is not None:
        dist.barrier()
    ms = trainer.timed_steps(args.steps)  # exactly K sweeps, CUDA events on the engine's stream, sync on both sides
    if dist is not None:
        import torch

        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop()
    launches = trainer.launch_count() - launches0
    # kernel families: a second, separately timed run with one CUDA-event pair per launch group (this run
    # launches directly; the timed run above replays the captured CUDA graph of the sweep)
    trainer.set_profiling(True)
    ms_profiled = trainer.timed_steps(args.steps)
    sweep_ms, sweep_launches = trainer.kernel_ms(0)
    qinit_ms, _ = trainer.kernel_ms(1)
    refresh_ms, _ = trainer.kernel_ms(2)
    stream_ms, _ = trainer.kernel_ms(3)
    gather_ms, _ = trainer.kernel_ms(4)
    trainer.set_profiling(False)
    hyper = trainer.get_hyper()
    sweep_path = trainer.sweep_path()
    del trainer

    it_per_s = args.steps / (ms / 1e3)
    real_bytes = 4 if dtype == "f32" else 8
    bytes_ = algorithmic_bytes(X.nnz, X.shape[0], rank, real_bytes)

    # ---- end-to-end leg: the call a user makes --------------------------------------------------
    stamps = []

    def callback(i, fm, hyper_, history):
        _ = (fm.w0, hyper_.alpha)  # the device->host read of the sweep's result
        stamps.append(time.perf_counter())
        return False, None

    if dist is not None:
        dist.barrier()
    t_fit0 = time.perf_counter()
    with options():
        model = myfm_b200.MyFMRegressor(rank=rank, random_seed=CHAIN_SEED)
        model.fit(X, y, n_iter=args.steps + args.warmup, n_kept_samples=1, group_shapes=group_shapes,
                  callback=callback)
    t_fit = time.perf_counter() - t_fit0
    e2e_s = stamps[-1] - stamps[args.warmup - 1] if args.warmup > 0 else stamps[-1] - t_fit0
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = args.steps / e2e_s
    G = len(group_shapes)
    # Host->device: a Gibbs fit uploads its inputs (X as CSR with int64 indptr / int32 indices / f64 values, y)
    # once, inside fit(); the sweeps read nothing else from the host (the variates come from the device-side
    # mt19937).  Reported amortised over the sweeps of this fit.  Device->host per sweep: the sweep's
    # hyper-parameters (LearningHistory) and the bias the callback reads.
    upload = X.indptr.shape[0] * 8 + X.nnz * (4 + 8) + y.shape[0] * 8
    h2d = upload / (args.steps + args.warmup)
    d2h = (2 + 2 * G + 2 * G * rank) * real_bytes + real_bytes

    if rank_id != 0:
        dist.destroy_process_group()
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = bytes_["sweeps"] * args.steps / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else None
    traffic, traffic_src = ncu_traffic(rank) if (sweep_path >= 1 and args.gpus == 1 and args.workload == "ml10m"
                                                 and dtype == "f32") else (None, None)
    line = {
        "metric": "gibbs_iterations_per_sec", "value": it_per_s, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": workload_config(args.workload, n_rows_global, nnz_global, rank, args.gpus),
        "nnz_rank_per_sec": it_per_s * nnz_global * rank,
        "e2e": {"value": e2e, "unit": "it/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "fit_total_s": t_fit, "fit_it_per_s_including_setup": (args.steps + args.warmup) / t_fit,
                "note": "MyFMRegressor.fit() with host scipy/numpy buffers; wall clock between per-iteration callbacks "
                        "(each reads the sweep's hyper-parameters and bias from the device).  The one-off input upload "
                        "(h2d_bytes_per_step = upload bytes / sweeps of this fit), transpose and level schedule are "
                        "inside fit_total_s"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "traffic_source": (f"{traffic_src}: dram read+write of k_field_stream + k_field_stats per launch x "
                                        f"{rank + 1} vectors per sweep (bytes per step, like algorithmic_bytes_per_step)")
                     if traffic else None, "peak_source": peak_src,
                     "kernel": ("k_field_stream + k_field_stats (column sweeps of w and of the K factor columns: "
                                "streaming level with fused q_init, gather-only last level)" if sweep_path == 1 else
                                "k_level_sweep + k_level_seg_update (column sweeps of w and of the K factor columns"
                                + ("; k_level_dist on this rank's row shard" if args.gpus > 1 else "") + ")"),
                     "algorithmic_bytes_per_step": bytes_["sweeps"], "launch_groups": int(sweep_launches),
                     "share_of_step": sweep_ms / ms_profiled if ms_profiled else None,
                     "whole_step_GBps": bytes_["total"] * it_per_s / 1e9,
                     "whole_step_frac": bytes_["total"] * it_per_s / 1e9 / peak},
        "kernel_ms_per_step": {"column_sweeps": sweep_ms / args.steps, "q_init": qinit_ms / args.steps,
                               "e_refresh": refresh_ms / args.steps, "streaming_level": stream_ms / args.steps,
                               "gather_level": gather_ms / args.steps,
                               "step_while_profiled": ms_profiled / args.steps},
        "alpha_last": hyper.alpha,
    }
    if args.gpus == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(X, y, group_shapes, rank)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ml10m", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
