"""The configurations of BASELINE.json as parity cases, through the public
API / C ABI against the oracle on the same seed:

C2  ML-100k-shaped, rank 8: held-out RMSE of the two predictors agrees to 1e-5
C3  ML-1M-extended-shaped (relation blocks with implicit-feedback features), rank 16: reduced rows
    (4 sweeps) and at the full block size of the configuration (2 sweeps)
C4  ML-10M-shaped at FULL size, rank 32, f32: two sweeps against the oracle (1e-4), plus
    size-independent properties: the residual cache equals prediction - y, runs are bit-reproducible
C5  64 categorical fields, ordered probit (reduced rows): 64 dependency levels, of which 62 run the
    general kernels between the field path's streaming and gather levels
"""
import numpy as np
import pytest
import scipy.sparse as sps

from helpers import fields_like, ml1m_extended, movielens_like
from test_gpu_parity import assert_state_close, close, make_pair, run_chain_parity

pytestmark = pytest.mark.gpu


def test_c2_heldout_rmse_matches_oracle(engine, oracle):
    X, y, gs = movielens_like(100_000, 943, 1682, 8, seed=0)
    y = np.clip(np.round(y), 1, 5)
    Xtr, ytr, Xte, yte = X[:80_000], y[:80_000], X[80_000:], y[80_000:]
    n_iter, n_kept = 30, 25
    for dtype, tol in (("f64", 1e-7), ("f32", 1e-5)):
        trainer, chain = make_pair(engine, oracle, Xtr, ytr, 8, dtype, group_shapes=gs, n_iter=n_iter)
        ours, theirs = np.zeros(Xte.shape[0]), np.zeros(Xte.shape[0])
        for it in range(n_iter):
            trainer.step(1)
            chain.step()
            if it >= n_iter - n_kept:
                w0, w, V, _ = trainer.get_fm()
                ours += oracle.predict_score(dtype, w0, w, V, Xte)
                ow0, ow, oV = chain.fm()
                theirs += oracle.predict_score(dtype, ow0, ow, oV, Xte)
        rmse = [float(np.sqrt(np.mean((p / n_kept - yte) ** 2))) for p in (ours, theirs)]
        assert abs(rmse[0] - rmse[1]) < tol * max(1.0, rmse[1]), (dtype, rmse)
        assert rmse[0] < 1.2  # and the model actually learned something (planted noise 0.9)


def test_c3_relation_blocks_rank16(engine, oracle):
    main, ub, mb, y, gs = ml1m_extended(30_000, 400, 250, 60, seed=1)
    trainer, chain = make_pair(engine, oracle, main, y, 16, "f64", X_rel=[ub, mb], group_shapes=gs)
    run_chain_parity(trainer, chain, "f64", 4)


def test_c3_full_block_size_two_sweeps(engine, oracle):
    """C3 at the size of BASELINE.json's configs[2] (bench.py --workload ml1m-ext): 900 188 rows, day one-hot
    main table, user block 6 040 x 9 746 and movie block 3 706 x 9 746 with implicit-feedback columns
    (one dependency level per implicit column), rank 16: two sweeps against the oracle, f64 at 1e-8."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    wl = bench.Workload("ml1m-ext")
    assert wl.rel[0][1].shape == (6040, 9746) and wl.rel[1][1].shape == (3706, 9746) and wl.n_rows == 900_188
    trainer, chain = make_pair(engine, oracle, wl.X, wl.y, wl.rank, "f64", X_rel=wl.rel, group_shapes=wl.group_shapes,
                               n_iter=4)
    run_chain_parity(trainer, chain, "f64", 2)


def test_c4_full_size_two_sweeps_and_properties(engine, oracle):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    X, y, gs, rank = bench.make_workload("ml10m")
    trainer, chain = make_pair(engine, oracle, X, y, rank, "f32", group_shapes=gs, n_iter=4)
    assert trainer.sweep_path() == 1  # the field path (csrc/field_sweep.cuh)
    twin, _unused = None, None
    for it in range(2):
        trainer.step(1)
        chain.step()
        assert_state_close(trainer, chain, "f32", f"full-size sweep {it}", free_running=(it > 0))
    # size-independent properties on the engine's own state
    w0, w, V, _ = trainer.get_fm()
    e = trainer.get_e()
    pred = oracle.predict_score("f64", w0, w, V, X[:200_000])
    np.testing.assert_allclose(e[:200_000], pred - y[:200_000], rtol=0, atol=2e-4)  # e == f(x) - y after update_e
    # bit-reproducibility: a second trainer on the same seed walks the same chain exactly
    from myfm_b200._myfm import ConfigBuilder, _TrainerHandle
    cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(gs)), gs)).set_n_iter(4)
           .set_n_kept_samples(4).build())
    with engine.engine_options(dtype="f32"):
        twin = _TrainerHandle(X, [], y, 42, cfg)
        twin.init_fm(rank, 0.1)
    twin.step(2)
    tw0, tw, tV, _ = twin.get_fm()
    assert tw0 == w0
    np.testing.assert_array_equal(tw, w)
    np.testing.assert_array_equal(tV, V)


def test_c5_many_fields_ordered_probit(engine, oracle):
    X, score, gs = fields_like(4000, [12] * 64, 4, seed=3, unit=True, noise=1.0)
    y = np.digitize(score, np.quantile(score, [0.2, 0.4, 0.6, 0.8])).astype(np.float64)
    trainer, chain = make_pair(engine, oracle, X, y, 6, "f64", task="ordered", group_shapes=gs)
    assert trainer.sweep_path() == 1
    run_chain_parity(trainer, chain, "f64", 3)
