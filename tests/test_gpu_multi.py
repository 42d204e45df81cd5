"""Row-sharded training on 2 GPUs (NCCL) against the single-GPU engine on the whole data.
Needs two visible devices: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import movielens_like

    # heavy tail: the top movie is a multi-chunk column on each shard
    return movielens_like(60001, 700, 25, 3, seed=11, zipf=(0.5, 1.2))


def _config(group_shapes, n_iter):
    from myfm_b200._myfm import ConfigBuilder

    return (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(group_shapes)), group_shapes))
            .set_n_iter(n_iter).set_n_kept_samples(n_iter).build())


def _state(trainer):
    w0, w, V, _ = trainer.get_fm()
    h = trainer.get_hyper()
    return np.concatenate([[w0], w, V.ravel(), [h.alpha], h.mu_w, h.lambda_w, h.mu_V.ravel(), h.lambda_V.ravel()])


def _worker(rank: int, world: int, port: int, dtype: str, n_sweeps: int, out_dir: str, exchange: str,
            no_field: bool, partition: str) -> None:
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ["MYFM_NO_PEER"] = "0" if exchange == "peer" else "1"
    os.environ["MYFM_NO_FIELD_PATH"] = "1" if no_field else "0"
    import torch
    import torch.distributed as dist

    import myfm_b200
    from myfm_b200 import distributed as mdist
    from myfm_b200._myfm import _TrainerHandle

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X, y, group_shapes = _data()
        X_local, y_local, ctx = mdist.shard(X, y, partition=partition)
        with ctx.options(dtype=dtype, device=rank):
            t = _TrainerHandle(X_local, [], y_local, 42, _config(group_shapes, n_sweeps))
            t.init_fm(6, 0.1)
        # 0 general level kernels, 1 field path + NCCL all-reduce, 2 field path + peer-memory exchange,
        # +2 when the first field's columns are rank-exclusive (rows dealt out by column)
        want = 0 if no_field else (2 if exchange == "peer" else 1) + (2 if partition == "column" else 0)
        assert t.sweep_path() == want, (t.sweep_path(), want)
        states = []
        for _ in range(n_sweeps):
            t.step(1)
            states.append(_state(t))
        np.save(os.path.join(out_dir, f"states_{rank}.npy"), np.stack(states))
        np.save(os.path.join(out_dir, f"e_{rank}.npy"), t.get_e())
        np.save(os.path.join(out_dir, f"rows_{rank}.npy"), ctx.rows)
        del t
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype,exchange,no_field,partition", [
    ("f64", "peer", False, "column"), ("f32", "peer", False, "column"), ("f64", "nccl", False, "column"),
    ("f64", "peer", False, "rows"), ("f32", "peer", False, "rows"), ("f64", "nccl", False, "rows"),
    ("f64", "nccl", True, "rows"), ("f64", "nccl", True, "column")])
def test_two_gpus_match_one(engine, dtype, exchange, no_field, partition, tmp_path):
    """2 row shards == 1 GPU on the whole data (f64 1e-8), replicas bit-identical, for every
    row-sharded schedule: rows dealt out by first column (the first field needs no exchange) or in
    contiguous ranges (two passes around the exchange); statistics exchanged through peer memory
    or NCCL all-reduces; field path or general level kernels."""
    from myfm_b200 import _lib

    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from myfm_b200._myfm import _TrainerHandle

    n_sweeps = 4
    mp.spawn(_worker, args=(2, _free_port(), dtype, n_sweeps, str(tmp_path), exchange, no_field, partition), nprocs=2,
             join=True)
    X, y, group_shapes = _data()
    with engine.engine_options(dtype=dtype):
        t = _TrainerHandle(X, [], y, 42, _config(group_shapes, n_sweeps))
        t.init_fm(6, 0.1)
    got = [np.load(tmp_path / f"states_{r}.npy") for r in range(2)]
    np.testing.assert_array_equal(got[0], got[1])  # replicas stay bit-identical
    tol = 1e-8 if dtype == "f64" else 2e-3
    for it in range(n_sweeps):
        t.step(1)
        want = _state(t)
        scale = np.maximum(np.abs(want), 1e-2)
        assert np.max(np.abs(got[0][it] - want) / scale) < tol, f"sweep {it}"
    e = np.empty(X.shape[0])
    for r in range(2):
        e[np.load(tmp_path / f"rows_{r}.npy")] = np.load(tmp_path / f"e_{r}.npy")
    np.testing.assert_allclose(e, t.get_e(), rtol=tol, atol=tol)


def _latent_data(task):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import fields_like

    # three fields: rank-exclusive first field, a middle level on the general kernels, a gather-only last level
    X, score, gs = fields_like(30011, [400, 30, 12], 3, seed=17, unit=True, noise=1.0)
    if task == "classification":
        y = np.where(score > np.median(score), 1.0, -1.0)
    else:
        y = np.digitize(score, np.quantile(score, [0.25, 0.5, 0.75])).astype(np.float64)
    return X, y, gs


def _latent_config(task, y, group_shapes, n_iter):
    from myfm_b200._myfm import ConfigBuilder, TaskType

    b = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(group_shapes)), group_shapes))
         .set_n_iter(n_iter).set_n_kept_samples(n_iter))
    b.set_task_type(TaskType.CLASSIFICATION if task == "classification" else TaskType.ORDERED)
    if task == "ordered":
        b.set_cutpoint_groups([(4, np.arange(y.shape[0]))])
    return b.build()


def _latent_worker(rank: int, world: int, port: int, task: str, n_sweeps: int, out_dir: str) -> None:
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    from myfm_b200 import distributed as mdist
    from myfm_b200._myfm import _TrainerHandle

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X, y, gs = _latent_data(task)
        X_local, y_local, ctx = mdist.shard(X, y, partition="column")
        with ctx.options(dtype="f64", device=rank, rng="philox"):
            t = _TrainerHandle(X_local, [], y_local, 42, _latent_config(task, y_local, gs, n_sweeps))
            t.init_fm(4, 0.1)
        states, cuts = [], []
        for _ in range(n_sweeps):
            t.step(1)
            states.append(_state(t))
            if task == "ordered":
                cuts.append(t.get_fm()[3][0])
        np.save(os.path.join(out_dir, f"states_{rank}.npy"), np.stack(states))
        if task == "ordered":
            np.save(os.path.join(out_dir, f"cuts_{rank}.npy"), np.stack(cuts))
            np.save(os.path.join(out_dir, f"accept_{rank}.npy"), np.asarray([t.mh_accept(0)]))
        del t
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("task", ["classification", "ordered"])
def test_two_gpus_latent_tasks_match_one(engine, task, tmp_path):
    """Row-sharded probit classification / ordered probit (rng="philox"): the latent draw of a row is keyed by
    its GLOBAL index and the cut-point sampler's row sums are all-reduced (OProbitSampler.hpp:389-463), so two
    shards walk the chain of one GPU on the whole data (f64 1e-8; cut-points and MH acceptances included)."""
    from myfm_b200 import _lib

    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from myfm_b200._myfm import _TrainerHandle

    n_sweeps = 4
    mp.spawn(_latent_worker, args=(2, _free_port(), task, n_sweeps, str(tmp_path)), nprocs=2, join=True)
    X, y, gs = _latent_data(task)
    with engine.engine_options(dtype="f64", rng="philox"):
        t = _TrainerHandle(X, [], y, 42, _latent_config(task, y, gs, n_sweeps))
        t.init_fm(4, 0.1)
    got = [np.load(tmp_path / f"states_{r}.npy") for r in range(2)]
    np.testing.assert_array_equal(got[0], got[1])
    for it in range(n_sweeps):
        t.step(1)
        want = _state(t)
        scale = np.maximum(np.abs(want), 1e-2)
        assert np.max(np.abs(got[0][it] - want) / scale) < 1e-8, f"sweep {it}"
        if task == "ordered":
            np.testing.assert_allclose(np.load(tmp_path / "cuts_0.npy")[it], t.get_fm()[3][0], rtol=1e-8, atol=1e-10)
    if task == "ordered":
        assert int(np.load(tmp_path / "accept_0.npy")[0]) == t.mh_accept(0)
