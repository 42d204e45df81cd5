"""The field path of the column sweeps (myfm_b200/csrc/field_sweep.cuh: streaming level 0 with the
fused q_init and the deferred last-level update, gather-only last level) against the oracle and
against the general dependency-level kernels, through the C ABI."""
import os

import numpy as np
import pytest

from helpers import fields_like, movielens_like
from test_gpu_parity import assert_state_close, make_pair, run_chain_parity

pytestmark = pytest.mark.gpu


def test_path_selection(engine, oracle, monkeypatch):
    X, y, gs = movielens_like(5000, 60, 20, 3, seed=1)
    t, _ = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=gs)
    assert t.sweep_path() == 1  # field path
    monkeypatch.setenv("MYFM_TILE_PATH", "1")
    t, _ = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=gs)
    assert t.sweep_path() == 5  # two fields, one GPU, opted in: tile path
    X3, y3, gs3 = fields_like(3000, [20, 10, 5], 2, seed=2)
    t, _ = make_pair(engine, oracle, X3, y3, 3, "f64", group_shapes=gs3)
    assert t.sweep_path() == 1  # three fields: field path even when opted in
    monkeypatch.delenv("MYFM_TILE_PATH")
    os.environ["MYFM_NO_FIELD_PATH"] = "1"
    try:
        t, _ = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=gs)
        assert t.sweep_path() == 0
    finally:
        del os.environ["MYFM_NO_FIELD_PATH"]
    # ragged rows (a row without its second field) are not a field stack
    Xr = X.tolil()
    Xr[0, X[0].indices[1]] = 0
    Xr = Xr.tocsr()
    Xr.eliminate_zeros()
    t, _ = make_pair(engine, oracle, Xr, y, 3, "f64", group_shapes=gs)
    assert t.sweep_path() == 0


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_long_columns(engine, oracle, dtype, two_field_path):
    """Heavy-tailed fields: level-0 columns above 1024 rows (whole CTA), 256..1024 (warp, two
    passes), below (registers); last-level columns above 8192 entries (chunked statistics).  On the
    tile path: a first-field column above 4096 rows (whole CTA), B-order segments of thousands of rows."""
    X, y, gs = movielens_like(60000, 200, 12, 4, seed=3, zipf=1.0)
    lens = np.diff(X.tocsc().indptr)
    assert lens[:200].max() > 4096 and lens[200:].max() > 8192 and lens[:200].min() < 256
    t, chain = make_pair(engine, oracle, X, y, 4, dtype, group_shapes=gs)
    assert t.sweep_path() == two_field_path
    run_chain_parity(t, chain, dtype, 6 if dtype == "f64" else 3)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_three_fields_with_values(engine, oracle, dtype):
    """Three fields, non-unit values: the middle level runs the general gather/scatter kernel
    between the streaming pass and the gather-only level."""
    X, y, gs = fields_like(8000, [50, 30, 20], 3, seed=5, unit=False)
    t, chain = make_pair(engine, oracle, X, y, 5, dtype, group_shapes=gs)
    assert t.sweep_path() == 1
    run_chain_parity(t, chain, dtype, 6 if dtype == "f64" else 3)


def test_four_unit_fields(engine, oracle):
    X, y, gs = fields_like(6000, [40, 25, 10, 7], 3, seed=6, unit=True)
    t, chain = make_pair(engine, oracle, X, y, 4, "f64", group_shapes=gs)
    assert t.sweep_path() == 1
    run_chain_parity(t, chain, "f64", 5)


@pytest.mark.parametrize("kw", [dict(fit_linear=False), dict(fit_w0=False), dict(fit_linear=False, fit_w0=False)])
def test_without_linear_or_bias(engine, oracle, kw, two_field_path):
    """fit_linear=False: the first factor's streaming pass finds nothing pending."""
    X, y, gs = movielens_like(8000, 80, 30, 3, seed=7)
    t, chain = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=gs, **kw)
    assert t.sweep_path() == two_field_path
    run_chain_parity(t, chain, "f64", 5)


def test_rank_zero_and_one(engine, oracle, two_field_path):
    X, y, gs = movielens_like(4000, 50, 20, 2, seed=8)
    for rank in (0, 1):
        t, chain = make_pair(engine, oracle, X, y, rank, "f64", group_shapes=gs)
        run_chain_parity(t, chain, "f64", 4)


@pytest.mark.parametrize("task", ["classification", "ordered"])
def test_latent_tasks(engine, oracle, task, two_field_path):
    X, y, gs = movielens_like(3000, 40, 15, 2, seed=9)
    if task == "classification":
        y = (y > np.median(y)).astype(np.float64) * 2 - 1
    else:
        y = np.digitize(y, np.quantile(y, [0.33, 0.66])).astype(np.float64)
    t, chain = make_pair(engine, oracle, X, y, 3, "f64", task=task, group_shapes=gs)
    assert t.sweep_path() == two_field_path
    run_chain_parity(t, chain, "f64", 4)


def test_field_path_equals_general_path(engine, oracle, two_field_path):
    """Same chain through both schedules: they differ in summation order only."""
    X, y, gs = fields_like(20000, [300, 60], 4, seed=10, unit=False)
    a, _ = make_pair(engine, oracle, X, y, 6, "f64", group_shapes=gs)
    os.environ["MYFM_NO_FIELD_PATH"] = "1"
    try:
        b, _ = make_pair(engine, oracle, X, y, 6, "f64", group_shapes=gs)
    finally:
        del os.environ["MYFM_NO_FIELD_PATH"]
    assert (a.sweep_path(), b.sweep_path()) == (two_field_path, 0)
    for it in range(5):
        a.step(1)
        b.step(1)
        for xa, xb in zip(a.get_fm()[:3], b.get_fm()[:3]):
            np.testing.assert_allclose(xa, xb, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(a.get_e(), b.get_e(), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(a.get_q(), b.get_q(), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("rank,unit,fields", [(20, True, [60, 25]), (40, False, [30, 20, 10]), (33, True, [40, 9, 5, 4])])
def test_wide_rank_forward_pass(engine, oracle, rank, unit, fields):
    """16 <= rank <= 64 with fixed-length rows: the tile forward pass (k_predict_tile, one or two
    factors per lane) refreshes e; also through the prediction-dataset entry point."""
    X, y, gs = fields_like(3000, fields, 3, seed=11, unit=unit)
    t, chain = make_pair(engine, oracle, X, y, rank, "f64", group_shapes=gs)
    run_chain_parity(t, chain, "f64", 3)
    from myfm_b200._myfm import FM

    w0, w, V, _ = t.get_fm()
    with engine.engine_options(dtype="f64"):
        score = FM(w0, w, V).predict_score(X[:1000], [])
    Xd = X[:1000]
    X2 = Xd.copy()
    X2.data = X2.data ** 2
    ref = w0 + Xd.dot(w) + 0.5 * ((Xd.dot(V) ** 2).sum(1) - X2.dot((V ** 2).sum(1)))
    np.testing.assert_allclose(score, ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("dtype", ["f32"])
def test_staged_streaming_kernel(engine, oracle, dtype, monkeypatch):
    """MYFM_STAGING=1: the warp class of k_field_stream reads its rows from per-warp shared-memory buffers filled
    by TMA bulk copies (cp.async.bulk + mbarrier) one column ahead; same chain as the direct-load kernel, bit for bit."""
    X, y, gs = movielens_like(40000, 900, 150, 4, seed=12, zipf=(0.5, 0.8))
    a, chain = make_pair(engine, oracle, X, y, 6, dtype, group_shapes=gs)
    monkeypatch.setenv("MYFM_STAGING", "1")
    b, _ = make_pair(engine, oracle, X, y, 6, dtype, group_shapes=gs)
    for it in range(4):
        a.step(1)
        b.step(1)
        chain.step()
        for xa, xb in zip(a.get_fm()[:3], b.get_fm()[:3]):
            np.testing.assert_array_equal(xa, xb)
        np.testing.assert_array_equal(a.get_e(), b.get_e())
    assert_state_close(b, chain, dtype, "staged kernel, sweep 3", free_running=True)
