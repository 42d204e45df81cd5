#!/usr/bin/env python
"""Headline benchmark: Gibbs iterations/s of the Bayesian-FM sweep at rank 32 on the
MovieLens-10M-shaped workload (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
    python bench.py --workload ml1m-ext                      # C3: relation blocks (rank 16)
    python bench.py --workload ml100k --task ordered         # probit / ordered-probit sweeps

One JSON line on stdout (rank 0).  A "step" is one full update_all sweep (alpha, w0, hypers, w,
K factor columns of V, e refresh) over the whole training set.

value     whole-job iterations/s, inputs resident in HBM, timed with CUDA events on the engine's
          stream around exactly K sweeps (barrier + synchronize on both sides, max over ranks).
e2e       the same metric through the public API (MyFMRegressor.fit with host scipy/numpy
          buffers): wall clock between the per-iteration callbacks (steady state), and the whole
          fit() including the one-off upload / data preparation (`including_setup`).
roofline  column-sweep kernels (the dominant family): algorithmic bytes / CUDA-event time.
At N > 1 the line also carries `parity_vs_1gpu`: the row-sharded chain against the 1-GPU chain on
the same seed after the same number of sweeps.
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
    # C3 (examples/ml-1m-extended.ipynb): day one-hot main table + user / movie relation blocks
    # carrying the id one-hot and SVD++-style implicit features
    "ml1m-ext": (900_188, 6_040, 3_706, 16),
}
ML1M_DAYS = 1_040
DATA_SEED = 2
CHAIN_SEED = 42
N_CLASSES = 5


class Workload:
    """Synthetic inputs of one configuration: main table, relation blocks (map, block), targets."""

    def __init__(self, name: str, task: str = "regression", rank: int | None = None):
        from helpers import ml1m_extended, movielens_like

        rows, users, movies, default_rank = WORKLOADS[name]
        self.name, self.task, self.rank = name, task, rank or default_rank
        if name == "ml1m-ext":
            X, ub, mb, y, gs = ml1m_extended(rows, users, movies, ML1M_DAYS, seed=DATA_SEED)
            self.X, self.rel, self.group_shapes = X, [ub, mb], gs
            y = np.clip(np.round(y), 1, 5)
        else:
            # popularity exponents matched to ML-10M's public max/mean ratings per user (7359/143)
            # and per movie (34864/937)
            X, y, gs = movielens_like(rows, users, movies, 8, seed=DATA_SEED, zipf=(0.4, 0.45))
            self.X, self.rel, self.group_shapes = X, [], gs
            y = np.clip(np.round(y * 2) / 2, 0.5, 5.0)  # half-star grid like ML-10M
        if task == "classification":
            y = (y > np.median(y)).astype(np.float64)  # 0 / 1 as MyFMClassifier takes it
        elif task == "ordered":
            y = np.clip(np.round(y), 1, N_CLASSES) - 1  # five ordinal classes 0 .. 4 (the star ratings)
        self.y = y
        self.nnz = int(X.nnz)
        self.n_rows = int(X.shape[0])

    @property
    def y_engine(self):
        """Targets as the trainer takes them (classification: -1 / +1, src/myfm/base.py:385-386)."""
        return self.y * 2 - 1 if self.task == "classification" else self.y

    def description(self) -> str:
        rows, users, movies, _ = WORKLOADS[self.name]
        if self.name == "ml1m-ext":
            nnz_b = [int(b.nnz) for _, b in self.rel]
            shape = (f"{rows} rows x {ML1M_DAYS} day one-hot + user block {users} x {users + movies} (nnz {nnz_b[0]}) "
                     f"+ movie block {movies} x {users + movies} (nnz {nnz_b[1]}), group_shapes={self.group_shapes}")
        else:
            shape = (f"{rows} rows x ({users} users + {movies} movies) one-hot, nnz={self.nnz}, "
                     f"group_shapes=[users, movies]")
        return f"{self.name}-shaped synthetic: {shape}, rank {self.rank}, {self.task}"

    def config(self, n_iter: int):
        from myfm_b200._myfm import ConfigBuilder, TaskType

        b = (ConfigBuilder().set_mu_0(0.0)
             .set_group_index(np.repeat(np.arange(len(self.group_shapes)), self.group_shapes))
             .set_n_iter(n_iter).set_n_kept_samples(1))
        b.set_task_type({"regression": TaskType.REGRESSION, "classification": TaskType.CLASSIFICATION,
                         "ordered": TaskType.ORDERED}[self.task])
        if self.task == "ordered":
            b.set_cutpoint_groups([(int(self.y.max()) + 1, np.arange(self.n_rows))])
        return b.build()

    def blocks(self):
        from myfm_b200._myfm import RelationBlock

        return [RelationBlock(m, b) for m, b in self.rel]

    def oracle_chain(self, dtype="f32"):
        from oracle import binding as oracle

        oracle.build()
        return oracle.OracleChain(self.X, self.y_engine, self.rank, X_rel=list(self.rel), dtype=dtype,
                                  task=self.task, seed=CHAIN_SEED, group_shapes=self.group_shapes, n_iter=10 ** 6)

    def algorithmic_bytes(self, real_bytes: int = 4):
        """bytes one sweep must move (i32 indices, `real_bytes` values)."""
        b, nnz, n, K = real_bytes, self.nnz, self.n_rows, self.rank
        v_sweep = (4 + b) * nnz + 4 * b * nnz          # per factor: CSC read + gather/scatter of q, e
        q_init = (4 + b) * nnz + 4 * n + b * n
        w_sweep = (4 + b) * nnz + 2 * b * nnz
        e_refresh = (4 + b) * nnz + 4 * n + 2 * b * n
        misc = 3 * b * n
        # relation blocks: per block per vector two passes over the row map with q, e
        # read + write, and the block sweep itself
        rel = sum((4 + 4 * b) * n * 2 + (4 + b) * int(B.nnz) for _, B in self.rel)
        sweeps = K * (v_sweep + rel) + w_sweep + rel
        return dict(sweeps=sweeps, total=sweeps + K * q_init + e_refresh + misc)


# ---- C5 (BASELINE.json configs[4]): 64 categorical fields x 31 250 categories, ordered probit, rank 64 ----
C5_FIELDS, C5_CATS, C5_RANK, C5_ROWS = 64, 31_250, 64, 50_000_000


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser: a counter-based hash, so any rank can produce any row's data."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def c5_field_column(rows: np.ndarray, field: int) -> np.ndarray:
    """Category (0 .. C5_CATS-1) of `field` for the global rows `rows`."""
    h = _mix64(rows.astype(np.uint64) * np.uint64(C5_FIELDS) + np.uint64(field) + np.uint64(DATA_SEED) * np.uint64(1 << 40))
    return (h % np.uint64(C5_CATS)).astype(np.int32)


class C5Shard:
    """This rank's rows of the C5 table, generated locally (no rank materialises the 3.2 G non-zeros):
    rows are dealt out by their first-field category (rank-exclusive first field), the other 63 fields,
    the planted score and the five ordinal classes come from counter-based hashes of the global row."""

    def __init__(self, n_rows: int, rank: int, world: int):
        from myfm_b200 import distributed as mdist

        self.n_rows_global, self.rank_id, self.world = n_rows, rank, world
        all_rows = np.arange(n_rows, dtype=np.int64)
        first = c5_field_column(all_rows, 0)
        self.rows = all_rows[mdist.partition_by_key(first, C5_CATS, world) == rank] if world > 1 else all_rows
        del all_rows, first
        n = self.rows.shape[0]
        cols = np.empty((n, C5_FIELDS), dtype=np.int32)
        rng = np.random.default_rng(DATA_SEED)
        w_planted = rng.normal(0, 0.25, C5_FIELDS * C5_CATS).astype(np.float32)
        score = np.zeros(n, dtype=np.float32)
        for f in range(C5_FIELDS):
            c = c5_field_column(self.rows, f)
            cols[:, f] = f * C5_CATS + c
            score += w_planted[f * C5_CATS + c]
        u = (_mix64(self.rows.astype(np.uint64) + np.uint64(1 << 50)) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
        v = (_mix64(self.rows.astype(np.uint64) + np.uint64(1 << 51)) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
        score = score + (np.sqrt(-2 * np.log(u + 1e-300)) * np.cos(2 * np.pi * v)).astype(np.float32)
        # the score is N(0, 64 * 0.25^2 + 1) = N(0, 5): quintiles of that normal
        cuts = np.sqrt(5.0) * np.asarray([-0.8416, -0.2533, 0.2533, 0.8416])
        self.y = np.digitize(score, cuts).astype(np.float64)
        import scipy.sparse as sps

        indptr = np.arange(0, C5_FIELDS * n + 1, C5_FIELDS, dtype=np.int64)
        self.X = sps.csr_matrix((np.ones(C5_FIELDS * n, dtype=np.float64), cols.ravel(), indptr),
                                shape=(n, C5_FIELDS * C5_CATS))
        self.group_shapes = [C5_CATS] * C5_FIELDS
        self.nnz_global = n_rows * C5_FIELDS

    def config(self, n_iter: int):
        from myfm_b200._myfm import ConfigBuilder, TaskType

        return (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(C5_FIELDS), C5_CATS))
                .set_n_iter(n_iter).set_n_kept_samples(1).set_task_type(TaskType.ORDERED)
                .set_cutpoint_groups([(N_CLASSES, np.arange(self.X.shape[0]))]).build())


def run_c5(args):
    """--workload c5: ordered probit on 64 fields x 31 250 categories, rank 64, rows sharded by their first
    field, Philox latent draws (statistical parity: SURVEY.md section 8d)."""
    import myfm_b200
    from myfm_b200._myfm import _TrainerHandle

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
    n_rows = args.rows or C5_ROWS
    t0 = time.perf_counter()
    shard = C5Shard(n_rows, rank_id, world)
    t_gen = time.perf_counter() - t0
    n_total = args.steps + args.warmup
    if dist is not None:
        from myfm_b200 import distributed as mdist

        ctx = mdist.context(shard.X, 0, n_rows)
        ctx.rows = shard.rows
        options = lambda: ctx.options(dtype=args.dtype, device=local, rng="philox")  # noqa: E731
    else:
        options = lambda: myfm_b200.engine_options(dtype=args.dtype, device=local, rng="philox")  # noqa: E731
    t0 = time.perf_counter()
    with options():
        trainer = _TrainerHandle(shard.X, [], shard.y, CHAIN_SEED, shard.config(n_total))
        trainer.init_fm(C5_RANK, 0.1)
    t_setup = time.perf_counter() - t0
    trainer.step(args.warmup)
    trainer.sync()
    launches0 = trainer.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    if dist is not None:
        dist.barrier()
    ms = trainer.timed_steps(args.steps)
    # end to end at the handle level: one sweep, then the sweep's hyper-parameters and cut-points to the host
    t0 = time.perf_counter()
    for _ in range(args.steps):
        trainer.step(1)
        hyper = trainer.get_hyper()
        cut = trainer.get_fm()[3][0] if False else None  # (the cut-points live on the host already)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        import torch

        t = torch.tensor([ms, e2e_s, t_setup, t_gen], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, t_setup, t_gen = (float(v) for v in t.tolist())
    clocks = sampler.stop()
    launches = trainer.launch_count() - launches0
    sweep_path = trainer.sweep_path()
    accept = trainer.mh_accept(0)
    cutpoints = trainer.get_fm()[3][0].tolist()
    del trainer
    if rank_id != 0:
        dist.destroy_process_group()
        return
    it_per_s = args.steps / (ms / 1e3)
    b = 4 if args.dtype == "f32" else 8
    nnz, K = n_rows * C5_FIELDS, C5_RANK
    # B_iter = K (32 nnz + 8 N) + 24 nnz + 24 N for f32 (whole job)
    v_sweep, q_init = (4 + b) * nnz + 4 * b * nnz, (4 + b) * nnz + 4 * n_rows + b * n_rows
    total_bytes = K * (v_sweep + q_init) + (4 + b) * nnz + 2 * b * nnz + (4 + b) * nnz + 4 * n_rows + 2 * b * n_rows
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    achieved = total_bytes * it_per_s / 1e9 / world
    path_name, kernel_desc = SWEEP_PATHS.get(sweep_path, (str(sweep_path), "?"))
    if wl.rel:
        kernel_desc += " + k_rel_sweep_smem (the relation blocks' column chain: one CTA, latency-bound, DESIGN.md 3d)"
    line = {
        "metric": "gibbs_iterations_per_sec", "value": it_per_s, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"c5-shaped synthetic: {n_rows} rows x {C5_FIELDS} categorical fields x {C5_CATS} "
                               f"categories (D = {C5_FIELDS * C5_CATS}), {C5_FIELDS} nnz per row (nnz = {nnz}), rank {K}, "
                               f"ordered probit with {N_CLASSES} classes, group_shapes=[{C5_CATS}] * {C5_FIELDS}; rows "
                               f"generated per shard from counter-based hashes",
                   "rank": K, "rows": n_rows, "nnz": nnz, "task": "ordered",
                   "rng": "philox latent draws on the device (statistical parity); Gaussian / Gamma variates from the "
                          "device-side mt19937 stream",
                   "parallelism": (f"rows dealt out over {world} GPUs by their first-field category (that field needs no "
                                   f"exchange); the other 63 dependency levels: column statistics all-reduced per level "
                                   f"(ncclAllReduce / peer memory); cut-point row sums all-reduced per Newton / MH evaluation")
                   if world > 1 else "single GPU"},
        "sweep_path": path_name, "nnz_rank_per_sec": it_per_s * nnz * K,
        "e2e": {"value": args.steps / e2e_s, "unit": "it/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int((2 + 2 * C5_FIELDS + 2 * C5_FIELDS * K) * b),
                "setup_s": t_setup, "generate_s": t_gen,
                "note": "trainer handle: one sweep, then the sweep's hyper-parameters on the host (what "
                        "create_train_fm does per iteration); the one-off upload / preparation is setup_s"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": kernel_desc + "; middle levels: k_level_dist / k_level_sweep",
                     "algorithmic_bytes_per_step": total_bytes, "note": "whole sweep, per GPU (SURVEY.md section 8d: "
                     "832 GB per iteration and GPU at 50 M rows on 8 GPUs)"},
        "cutpoints_last": cutpoints, "mh_accept": int(accept), "alpha_last": hyper.alpha,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


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


def cpu_baseline(wl: Workload, budget_s: float = 20.0):
    """The reference's CPU path (oracle port, f32, 1 thread) on the same workload; bounded."""
    chain = wl.oracle_chain()
    t1 = chain.timed_steps(1)  # warm-up, also sizes the sample
    n = int(max(1, min(20, budget_s // max(t1, 1e-3))))
    t = chain.timed_steps(n)
    return {"value": n / t, "unit": "it/s", "cores": 1, "kind": "port",
            "sample": f"{n} full update_all sweeps (after 1 warm-up) of the same workload, oracle restatement of the "
                      f"reference sampler (Real=float, g++ -O3, single thread as the reference is; host has "
                      f"{os.cpu_count()} cores)"}


# DRAM bytes per launch of the column-sweep kernels from the newest committed `ncu --set full` capture
SWEEP_KERNELS = ("k_field_stream", "k_field_stats", "k_tile_sweep", "k_tile_fold")


def ncu_traffic(rank: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the column-sweep kernels from the newest
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
        for key in SWEEP_KERNELS:
            if key in r[i_name]:
                b = float(r[i_rd]) * scale.get(units[i_rd], 1.0) + float(r[i_wr]) * scale.get(units[i_wr], 1.0)
                per_kernel.setdefault(key, []).append(b)
    if not per_kernel:
        return None, None
    per_vector = sum(sum(v) / len(v) for v in per_kernel.values())
    return per_vector * (rank + 1), f"{os.path.relpath(files[-1], ROOT)} ({' + '.join(sorted(per_kernel))})"


def base_config(wl: Workload, n_gpus: int, rng: str, parallelism: str):
    return {"workload": wl.description(), "rank": wl.rank, "rows": wl.n_rows, "nnz": wl.nnz, "task": wl.task,
            "rng": ("mt19937 (reference stream, same seed; generated on the device)" if rng == "mt19937" else
                    "philox latent draws on the device (statistically equivalent chain); Gaussian / Gamma variates "
                    "from the device-side mt19937 stream"),
            "l2": "working set (CSR+CSC+{e,q},y) exceeds the 126 MB L2 for ml10m; no explicit flush",
            "parallelism": parallelism}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle port; the reference cannot be
    built without Eigen) on the host cores it can use (one: the sampler is single-threaded)."""
    rank_id = int(os.environ.get("RANK", "0"))
    if rank_id != 0:
        return
    wl = Workload(args.workload, args.task, args.rank)
    chain = wl.oracle_chain()
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
        "config": base_config(wl, args.gpus, "mt19937", "CPU, 1 thread"),
        "nnz_rank_per_sec": value * wl.nnz * wl.rank,
        "cpu_baseline": {"value": value, "unit": "it/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} full update_all sweeps (of {args.steps} asked; bounded to ~3 min) of the same "
                                   f"workload after {1 + warm} warm-up, oracle restatement of the reference sampler "
                                   f"(Real=float, g++ -O3), 1 thread: the reference sampler is single-threaded "
                                   f"(host has {os.cpu_count()} cores)"},
        "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# myfm_trainer_sweep_path -> what runs
SWEEP_PATHS = {
    0: ("general", "k_level_sweep + k_level_seg_update (dependency-level column sweeps, gather + scatter)"),
    1: ("field", "k_field_stream + k_field_stats (column sweeps of w and of the K factor columns: streaming level "
                 "with fused q_init, gather-only last level)"),
    2: ("field+peer", "k_field_stream<STATS/UPDATE> + k_field_stats + k_field_draw_last, column statistics "
                      "exchanged through peer memory inside the kernels"),
    3: ("field-exclusive+nccl", "k_field_stream (rank-exclusive first field: no exchange) + k_field_stats + "
                                "ncclAllReduce + k_field_draw_last"),
    4: ("field-exclusive+peer", "k_field_stream (rank-exclusive first field: no exchange) + k_field_stats + "
                                "k_field_draw_last reading the peers' statistics over NVLink"),
    5: ("tile", "k_tile_sweep + k_tile_fold (row tiles staged in shared memory by TMA bulk copies: pending update, "
                "first-field sweep and last-field statistics in one pass)"),
    6: ("tile+peer", "k_tile_sweep + k_tile_fold (rank-exclusive row tiles; per-column statistics of the last "
                     "field exchanged through peer memory)"),
}


def parallelism_label(n_gpus: int, sweep_path: int, partition: str) -> str:
    if n_gpus == 1:
        return "single GPU"
    how = {0: "one ncclAllReduce of the column statistics per dependency level",
           1: "ncclAllReduce of the column statistics per level (field path)",
           2: "column statistics of both levels exchanged through peer memory (NVLink) inside the sweep kernels",
           3: "first field rank-exclusive (no exchange), last field: ncclAllReduce of its column statistics per vector",
           4: "first field rank-exclusive (no exchange), last field: column statistics read from the peers over "
              "NVLink inside the draw kernel, one exchange per vector",
           6: "row tiles rank-exclusive by first field; last field: per-column statistics read from the peers over "
              "NVLink, one exchange per vector"}.get(sweep_path, "?")
    return (f"rows sharded over {n_gpus} GPUs (partition='{partition}': rows dealt out by their first-field column), "
            f"model replicated; {how}")


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b)) / scale)


def run_ours(args):
    import myfm_b200
    from myfm_b200._myfm import _TrainerHandle

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

    wl = Workload(args.workload, args.task, args.rank)
    rank = wl.rank
    X, y = wl.X, wl.y_engine
    dtype = args.dtype
    n_total = args.steps + args.warmup
    cfg = wl.config(n_total)
    partition = "column"
    if dist is not None:  # every rank generated the same data; keep this rank's rows
        if wl.rel:
            raise SystemExit("relation blocks are single-GPU (the blocks' row caches are not sharded yet)")
        from myfm_b200 import distributed as mdist

        X, y, ctx = mdist.shard(X, y, partition=partition)
        options = lambda: ctx.options(dtype=dtype, device=local, rng=args.rng)  # noqa: E731
    else:
        options = lambda: myfm_b200.engine_options(dtype=dtype, device=local, rng=args.rng)  # noqa: E731

    # ---- device-resident leg ---------------------------------------------------------------
    with options():
        trainer = _TrainerHandle(X, wl.blocks(), y, CHAIN_SEED, cfg)
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

    # ---- N > 1: the sharded chain against the 1-GPU chain, same seed, same number of sweeps ----
    parity = None
    if dist is not None:
        n_par = 3
        with options():
            tp = _TrainerHandle(X, [], y, CHAIN_SEED, cfg)
            tp.init_fm(rank, 0.1)
        tp.step(n_par)
        tp.sync()
        w0s, ws, Vs, _ = tp.get_fm()
        hs = tp.get_hyper()
        del tp
        dist.barrier()
        if rank_id == 0:
            with myfm_b200.engine_options(dtype=dtype, device=local, rng=args.rng):
                t1 = _TrainerHandle(wl.X, [], wl.y_engine, CHAIN_SEED, cfg)
                t1.init_fm(rank, 0.1)
            t1.step(n_par)
            t1.sync()
            w01, w1, V1, _ = t1.get_fm()
            h1 = t1.get_hyper()
            del t1
            parity = {"sweeps": n_par, "max_rel_w": rel_err(ws, w1), "max_rel_V": rel_err(Vs, V1),
                      "rel_w0": abs(w0s - w01) / max(abs(w01), 1e-30), "rel_alpha": abs(hs.alpha - h1.alpha) / h1.alpha,
                      "alpha_equal": bool(hs.alpha == h1.alpha),
                      "note": "max |sharded - single| / max |single| per array after `sweeps` sweeps from the same seed "
                              f"({dtype}: the cross-rank sums add the same numbers in another order)"}
        dist.barrier()

    it_per_s = args.steps / (ms / 1e3)
    real_bytes = 4 if dtype == "f32" else 8
    bytes_ = wl.algorithmic_bytes(real_bytes)

    # ---- end-to-end leg: the call a user makes --------------------------------------------------
    stamps = []

    def callback(i, fm, hyper_, history):
        _ = hyper_.alpha  # the device->host read of the sweep's result: its hyper-parameters (fetched before this call)
        stamps.append(time.perf_counter())
        return False, None

    callback.observer = True  # reads `hyper` only: the next sweep may start before this host code runs

    if dist is not None:
        dist.barrier()
    t_fit0 = time.perf_counter()
    with options():
        if wl.task == "regression":
            model = myfm_b200.MyFMRegressor(rank=rank, random_seed=CHAIN_SEED)
        elif wl.task == "classification":
            model = myfm_b200.MyFMClassifier(rank=rank, random_seed=CHAIN_SEED)
        else:
            model = myfm_b200.MyFMOrderedProbit(rank=rank, random_seed=CHAIN_SEED)
        # sharded runs hand fit() this rank's rows; targets in the estimator's own convention
        y_fit = wl.y if dist is None else np.asarray(wl.y)[ctx.rows]
        model.fit(X, y_fit, X_rel=wl.blocks(), n_iter=n_total, n_kept_samples=1, group_shapes=wl.group_shapes,
                  callback=callback)
    t_fit = time.perf_counter() - t_fit0
    e2e_s = stamps[-1] - stamps[args.warmup - 1] if args.warmup > 0 else stamps[-1] - t_fit0
    setup_s = stamps[0] - t_fit0  # everything before the first sweep's callback, minus one sweep below
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s, t_fit, setup_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, t_fit, setup_s = (float(v) for v in t.tolist())
    e2e = args.steps / e2e_s
    setup_s = max(0.0, setup_s - e2e_s / args.steps)
    G = len(wl.group_shapes)
    # Host->device: a Gibbs fit uploads its inputs (X as CSR with int64 indptr / int32 indices / f64 values, y)
    # once, inside fit(); the sweeps read nothing else from the host (the variates come from the device-side
    # mt19937).  Reported amortised over the sweeps of this fit.  Device->host per sweep: the sweep's
    # hyper-parameters (LearningHistory).
    upload = X.indptr.shape[0] * 8 + X.nnz * (4 + 8) + y.shape[0] * 8
    upload += sum(B.indptr.shape[0] * 8 + B.nnz * 12 + np.asarray(m).shape[0] * 8 for m, B in wl.rel)
    h2d = upload / n_total
    d2h = (2 + 2 * G + 2 * G * rank) * real_bytes
    if wl.task != "regression" and args.rng == "mt19937":  # the latent draws run on the host: e both ways
        h2d += X.shape[0] * real_bytes
        d2h += X.shape[0] * real_bytes

    if rank_id != 0:
        dist.destroy_process_group()
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per GPU: every rank sweeps 1/N of the rows, and `peak` is one GPU's
    achieved = bytes_["sweeps"] / args.gpus * args.steps / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else None
    traffic, traffic_src = ncu_traffic(rank) if (sweep_path in (1, 5) and args.gpus == 1 and args.workload == "ml10m"
                                                 and dtype == "f32") else (None, None)
    path_name, kernel_desc = SWEEP_PATHS.get(sweep_path, (str(sweep_path), "?"))
    if wl.rel:
        kernel_desc += " + k_rel_sweep_smem (the relation blocks' column chain: one CTA, latency-bound, DESIGN.md 3d)"
    config_n_iter = 200  # the iteration count BASELINE.json's C4 is quoted on
    line = {
        "metric": "gibbs_iterations_per_sec", "value": it_per_s, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": base_config(wl, args.gpus, args.rng, parallelism_label(args.gpus, sweep_path, partition)),
        "sweep_path": path_name,
        "nnz_rank_per_sec": it_per_s * wl.nnz * rank,
        "e2e": {"value": e2e, "unit": "it/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "fit_total_s": t_fit, "setup_s": setup_s,
                "including_setup": {"it_per_s_this_fit": n_total / t_fit, "n_iter_this_fit": n_total,
                                    "it_per_s_at_config_n_iter": config_n_iter / (setup_s + config_n_iter / e2e),
                                    "config_n_iter": config_n_iter},
                "note": "value: MyFM*.fit() with host scipy/numpy buffers, wall clock between per-iteration callbacks "
                        "(each reads the sweep's hyper-parameters from the device; an `observer` callback, so the next sweep is already running while it executes — what fit() does with its own progress callback), steady state.  "
                        "including_setup: the whole fit() with its one-off input upload (h2d_bytes_per_step = upload "
                        "bytes / sweeps of this fit), transposes, level schedule; measured for this fit's n_iter and "
                        "projected (setup_s + n / value) to the configuration's own n_iter"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "traffic_source": (f"{traffic_src}: dram read+write per launch x {rank + 1} vectors per sweep "
                                        f"(bytes per step, like algorithmic_bytes_per_step)") if traffic else None,
                     "peak_source": peak_src, "kernel": kernel_desc,
                     "algorithmic_bytes_per_step": bytes_["sweeps"], "per_gpu": True, "launch_groups": int(sweep_launches),
                     "share_of_step": sweep_ms / ms_profiled if ms_profiled else None,
                     "frac_of_measured_traffic": (traffic * args.steps / (sweep_ms / 1e3) / 1e9 / peak)
                     if traffic and sweep_ms > 0 else None,
                     "whole_step_GBps": bytes_["total"] / args.gpus * it_per_s / 1e9,
                     "whole_step_frac": bytes_["total"] / args.gpus * it_per_s / 1e9 / peak},
        "kernel_ms_per_step": {"column_sweeps": sweep_ms / args.steps, "q_init": qinit_ms / args.steps,
                               "e_refresh": refresh_ms / args.steps, "streaming_level": stream_ms / args.steps,
                               "gather_level": gather_ms / args.steps,
                               "step_while_profiled": ms_profiled / args.steps},
        "alpha_last": hyper.alpha,
    }
    if parity is not None:
        line["parity_vs_1gpu"] = parity
    if args.gpus == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ml10m", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--rows", type=int, default=None, help="c5: total training rows (default 50 000 000)")
    ap.add_argument("--task", default="regression", choices=["regression", "classification", "ordered"])
    ap.add_argument("--rank", type=int, default=None, help="override the workload's rank")
    ap.add_argument("--rng", default="mt19937", choices=["mt19937", "philox"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "c5":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "c5 is sized for 8 GPUs; the CPU oracle is not run on it"}))
        else:
            run_c5(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


def make_workload(name: str):
    """(X, y, group_shapes, rank) of a two-field workload (tests and tools use this)."""
    wl = Workload(name)
    return wl.X, wl.y, wl.group_shapes, wl.rank


if __name__ == "__main__":
    main()
