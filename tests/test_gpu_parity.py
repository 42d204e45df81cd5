"""Parity of the CUDA engine with the oracle (oracle/fm_oracle.hpp) on the same seeded inputs.
Every call goes through the C ABI (include/myfm_b200.h) via myfm_b200._myfm.

Tolerances (BASELINE.json north_star): posterior samples within 1e-4 relative of the CPU path on
the same seed in f32; the f64 engine tracks the f64 oracle to ~1e-9 because the only differences
are summation order (warp trees vs a serial accumulator) and the last bit of log/sqrt.
"""
import numpy as np
import pytest
import scipy.sparse as sps

from helpers import block_data, dense_block_data, middle_data, movielens_like, toy_matrix

pytestmark = pytest.mark.gpu

# rtol is the north-star bar; atol is the same bar relative to the magnitude of the vector being
# compared (entries much smaller than the vector's scale carry the absolute rounding noise of the
# sums they came from), applied in assert_state_close.
TOL = {"f64": 1e-8, "f32": 1e-4}


FREE_RUNNING_F32 = 2e-3


def close(actual, desired, dtype, what, tol=None):
    desired = np.asarray(desired)
    tol = TOL[dtype] if tol is None else tol
    scale = max(1e-2, float(np.max(np.abs(desired)))) if desired.size else 1.0
    np.testing.assert_allclose(actual, desired, rtol=tol, atol=tol * scale, err_msg=what)


def make_pair(engine, oracle, X, y, rank, dtype, task="regression", X_rel=(), group_shapes=None,
              seed=42, n_iter=20, **kw):
    """Returns (engine trainer handle, oracle chain) initialised identically."""
    from myfm_b200._myfm import ConfigBuilder, RelationBlock, TaskType, _TrainerHandle

    blocks = [RelationBlock(m, b) for m, b in X_rel]
    dim_all = X.shape[1] + sum(b.feature_size for b in blocks)
    builder = ConfigBuilder()
    for key in ("alpha_0", "beta_0", "gamma_0", "reg_0"):
        getattr(builder, "set_" + key)(kw.get(key, 1.0))
    builder.set_mu_0(kw.get("mu_0", 0.0))
    builder.set_fit_w0(kw.get("fit_w0", True)).set_fit_linear(kw.get("fit_linear", True))
    if group_shapes is None:
        builder.set_identical_groups(dim_all)
    else:
        builder.set_group_index(np.repeat(np.arange(len(group_shapes)), group_shapes))
    builder.set_n_iter(n_iter).set_n_kept_samples(n_iter)
    builder.set_task_type({"regression": TaskType.REGRESSION, "classification": TaskType.CLASSIFICATION,
                           "ordered": TaskType.ORDERED}[task])
    if task == "ordered":
        builder.set_cutpoint_groups([(int(y.max()) + 1, np.arange(y.shape[0]))])
    with engine.engine_options(dtype=dtype):
        trainer = _TrainerHandle(X, blocks, y, seed, builder.build())
        trainer.init_fm(rank, 0.1)
    chain = oracle.OracleChain(X, y, rank, X_rel=list(X_rel), dtype=dtype, task=task, seed=seed,
                               group_shapes=group_shapes, n_iter=n_iter, n_kept_samples=n_iter,
                               mu_0=kw.get("mu_0", 0.0), fit_w0=kw.get("fit_w0", True),
                               fit_linear=kw.get("fit_linear", True),
                               **{k: kw[k] for k in ("alpha_0", "beta_0", "gamma_0", "reg_0") if k in kw})
    return trainer, chain


def assert_state_close(trainer, chain, dtype, what="", free_running=False):
    tol = FREE_RUNNING_F32 if (free_running and dtype == "f32") else None
    w0, w, V, _ = trainer.get_fm()
    ow0, ow, oV = chain.fm()
    close(w0, ow0, dtype, f"w0 {what}", tol)
    close(w, ow, dtype, f"w {what}", tol)
    close(V, oV, dtype, f"V {what}", tol)
    hyper, oh = trainer.get_hyper(), chain.hyper()
    for key in ("alpha", "mu_w", "lambda_w", "mu_V", "lambda_V"):
        close(getattr(hyper, key), oh[key], dtype, f"{key} {what}", tol)
    close(trainer.get_e(), chain.e(), dtype, f"e {what}", tol)


def run_chain_parity(trainer, chain, dtype, n_steps):
    """Free-running chains on the same seed.  f64 holds 1e-8 throughout.  In f32 the two
    implementations differ by summation order only, but the chain itself amplifies rounding noise
    from sweep to sweep (a different summation order on the CPU does the same), so free-running
    f32 chains are compared at 2e-3; the 1e-4 bar is enforced per sweep by
    test_teacher_forced_f32_sweeps."""
    assert_state_close(trainer, chain, dtype, "after init")
    for it in range(n_steps):
        trainer.step(1)
        chain.step()
        assert_state_close(trainer, chain, dtype, f"after sweep {it}", free_running=True)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_toy_c1_regression(engine, oracle, dtype):
    """Config C1: the 4 x 9 README matrix, rank 4 (one dense column + two one-hot fields: three
    dependency levels)."""
    X, y = toy_matrix()
    trainer, chain = make_pair(engine, oracle, X, y, 4, dtype)
    run_chain_parity(trainer, chain, dtype, 10)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_planted_dense_columns(engine, oracle, dtype):
    """Reference fixture (tests/conftest.py:26-45): 3 dense-ish columns -> 3 serial levels of one
    long column each (block-per-column kernel)."""
    X, score = middle_data(3000)
    y = score + np.random.RandomState(0).normal(0, 1, size=score.shape)
    trainer, chain = make_pair(engine, oracle, X, y, 3, dtype)
    run_chain_parity(trainer, chain, dtype, 10)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_movielens_shaped(engine, oracle, dtype):
    """Config C2 in miniature: user + movie one-hot, two groups, power-law column lengths that
    exercise both the warp-per-column and the block-per-column kernels."""
    X, y, group_shapes = movielens_like(20000, 300, 40, 4, seed=0)
    trainer, chain = make_pair(engine, oracle, X, y, 8, dtype, group_shapes=group_shapes)
    # free-running f32 chains drift apart at the rate the chain amplifies rounding noise; the
    # long f32 comparison is the teacher-forced test below
    run_chain_parity(trainer, chain, dtype, 8 if dtype == "f64" else 4)


@pytest.mark.parametrize("task", ["regression", "classification"])
def test_teacher_forced_f32_sweeps(engine, oracle, task):
    """f32, 20 sweeps, each started from the ORACLE's state: isolates the error one sweep of the
    device arithmetic adds (summation order, logf/sqrtf) from the chain's own sensitivity to
    rounding, which any two f32 implementations with different summation order show."""
    X, y, group_shapes = movielens_like(20000, 300, 40, 4, seed=0)
    if task == "classification":
        y = (y > np.median(y)).astype(np.float64) * 2 - 1
    trainer, chain = make_pair(engine, oracle, X, y, 8, "f32", task=task, group_shapes=group_shapes)
    from myfm_b200._myfm import FMHyperParameters

    for it in range(20):
        w0, w, V = chain.fm()
        h = chain.hyper()
        trainer.set_state(w0, w, V, FMHyperParameters(h["alpha"], h["mu_w"], h["lambda_w"], h["mu_V"],
                                                      h["lambda_V"]), chain.e())
        trainer.step(1)
        chain.step()
        assert_state_close(trainer, chain, "f32", f"teacher-forced sweep {it}")


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_heavy_tail_columns(engine, oracle, dtype):
    """One movie holds ~30 % of the rows: its column is cut into many multi-block segments."""
    X, y, group_shapes = movielens_like(30000, 500, 12, 3, seed=5, zipf=(0.5, 1.5))
    trainer, chain = make_pair(engine, oracle, X, y, 4, dtype, group_shapes=group_shapes)
    run_chain_parity(trainer, chain, dtype, 4)


@pytest.mark.parametrize("fit_w0,fit_linear", [(False, True), (True, False), (False, False)])
def test_fit_flags(engine, oracle, fit_w0, fit_linear):
    """FMTrainer.hpp:219-222,232-235: the zeroed parameter's stale contribution stays in e until
    update_e — and the RNG stream skips the corresponding draws."""
    X, y, group_shapes = movielens_like(3000, 60, 25, 3, seed=1)
    trainer, chain = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=group_shapes,
                               fit_w0=fit_w0, fit_linear=fit_linear)
    run_chain_parity(trainer, chain, "f64", 5)


@pytest.mark.parametrize("maker", [block_data, dense_block_data])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_relation_blocks_vs_oracle(engine, oracle, maker, dtype):
    X_flat, main, users, items, y, group_shapes = maker()
    trainer, chain = make_pair(engine, oracle, main, y, 2, dtype, X_rel=[users, items],
                               group_shapes=group_shapes, fit_w0=False)
    run_chain_parity(trainer, chain, dtype, 8)


@pytest.mark.parametrize("maker", [block_data, dense_block_data])
def test_relation_blocks_equal_flat_on_device(engine, oracle, maker):
    """The reference's strongest identity (tests/regression/test_block.py:136-139): every kept
    sample's V of the blocked model == the flattened model, rtol 1e-7, f64, 30 sweeps."""
    X_flat, main, users, items, y, group_shapes = maker()
    flat, _ = make_pair(engine, oracle, X_flat, y, 2, "f64", group_shapes=group_shapes, fit_w0=False)
    blocked, _ = make_pair(engine, oracle, main, y, 2, "f64", X_rel=[users, items],
                           group_shapes=group_shapes, fit_w0=False)
    for _ in range(30):
        flat.step(1)
        blocked.step(1)
        np.testing.assert_allclose(flat.get_fm()[2], blocked.get_fm()[2], rtol=1e-7)
        np.testing.assert_allclose(flat.get_fm()[1], blocked.get_fm()[1], rtol=1e-7)


def test_only_relation_blocks(engine, oracle):
    """X=None: a main table without columns (reference base.py:230-233)."""
    _, _, users, items, y, group_shapes = block_data()
    empty = sps.csr_matrix((y.shape[0], 0), dtype=np.float64)
    trainer, chain = make_pair(engine, oracle, empty, y, 2, "f64", X_rel=[users, items],
                               group_shapes=group_shapes[1:])
    run_chain_parity(trainer, chain, "f64", 5)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_classification_chain(engine, oracle, dtype):
    X, score = middle_data(1500)
    noised = score + np.random.RandomState(0).normal(0, 1, size=score.shape)
    y = ((noised - noised.mean()) > 0).astype(np.float64) * 2 - 1
    trainer, chain = make_pair(engine, oracle, X, y, 3, dtype, task="classification")
    run_chain_parity(trainer, chain, dtype, 8 if dtype == "f64" else 3)


def test_rank_zero_and_empty_columns(engine, oracle):
    """rank 0 (ordered-probit test of the reference uses it) and columns without entries."""
    X, y, _ = movielens_like(500, 30, 10, 2, seed=4)
    X = sps.hstack([X, sps.csr_matrix((500, 3))]).tocsr()
    trainer, chain = make_pair(engine, oracle, X, y, 0, "f64")
    run_chain_parity(trainer, chain, "f64", 4)
    trainer, chain = make_pair(engine, oracle, X, y, 2, "f64")
    run_chain_parity(trainer, chain, "f64", 4)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_predict_score_parity(engine, oracle, dtype):
    from myfm_b200._myfm import FM, RelationBlock

    X_flat, main, users, items, y, _ = dense_block_data()
    rng = np.random.default_rng(0)
    D = X_flat.shape[1]
    w0, w, V = 0.3, rng.normal(size=D), rng.normal(size=(D, 5)) * 0.3
    blocks = [RelationBlock(*users), RelationBlock(*items)]
    with engine.engine_options(dtype=dtype):
        got_flat = FM(w0, w, V).predict_score(X_flat, [])
        got_blk = FM(w0, w, V).predict_score(main, blocks)
    want = oracle.predict_score(dtype, w0, w, V, X_flat)
    tol = dict(rtol=1e-10, atol=1e-10) if dtype == "f64" else dict(rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got_flat, want, **tol)
    np.testing.assert_allclose(got_blk, want, **tol)
    np.testing.assert_allclose(oracle.predict_score(dtype, w0, w, V, main, [users, items]), want, **tol)


def test_golden_toy_chain(engine):
    """Committed golden vectors (tests/golden/make_golden.py ran the oracle): config C1, f64."""
    import os

    from myfm_b200._myfm import ConfigBuilder, _TrainerHandle

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "toy_c1_f64.npz"))
    X, y = toy_matrix()
    cfg = ConfigBuilder().set_mu_0(0.0).set_identical_groups(9).set_n_iter(10).set_n_kept_samples(10).build()
    with engine.engine_options(dtype="f64"):
        trainer = _TrainerHandle(X, [], y, 42, cfg)
        trainer.init_fm(4, 0.1)
    for it in range(10):
        trainer.step(1)
        w0, w, V, _ = trainer.get_fm()
        np.testing.assert_allclose(w0, g["w0"][it], rtol=1e-8)
        np.testing.assert_allclose(w, g["w"][it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(V, g["V"][it], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(trainer.get_hyper().alpha, g["alpha"][it], rtol=1e-8)


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("rank", [0, 2, 8])
def test_device_mt19937_stream_matches_libstdcxx(engine, dtype, rank, monkeypatch):
    """Regression sweeps draw their variates from a std::mt19937 stream reproduced ON THE DEVICE
    (csrc/mt_device.cuh).  The same chain with MYFM_HOST_RNG=1 draws them with libstdc++ itself:
    the two variate vectors must agree position by position (acceptance decisions exact; values
    to the last ulp or two of log) over several sweeps, which also checks the generator hand-over
    between sweeps."""
    from myfm_b200._myfm import ConfigBuilder, _TrainerHandle

    X, y, group_shapes = movielens_like(4000, 700, 300, 3, seed=3)
    cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat([0, 1], group_shapes))
           .set_n_iter(5).set_n_kept_samples(5).build())

    def variates(host: bool):
        if host:
            monkeypatch.setenv("MYFM_HOST_RNG", "1")
        else:
            monkeypatch.delenv("MYFM_HOST_RNG", raising=False)
        with engine.engine_options(dtype=dtype):
            t = _TrainerHandle(X, [], y, 7, cfg)
            t.init_fm(rank, 0.1)
        out = []
        for _ in range(4):
            t.step(1)
            t.sync()
            out.append(t.get_variates())
        return out

    dev, host = variates(False), variates(True)
    tol = 1e-6 if dtype == "f32" else 1e-14
    for it, (a, b) in enumerate(zip(dev, host)):
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=tol, atol=0, err_msg=f"sweep {it}")


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_device_mt19937_farm_matches_libstdcxx(engine, dtype, monkeypatch):
    """The parallel word generator (k_mt_farm: several lanes, each jumping over the words of the
    others with the GF(2) jump polynomial of csrc/mt_jump.hpp) against libstdc++ draw for draw,
    across several block boundaries: 4 lanes x 32768 words per block, ~85 k (f32) / ~170 k (f64)
    words per sweep."""
    from myfm_b200._myfm import ConfigBuilder, _TrainerHandle

    X, y, group_shapes = movielens_like(4000, 700, 300, 3, seed=3)
    cfg = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat([0, 1], group_shapes))
           .set_n_iter(8).set_n_kept_samples(8).build())

    def variates(host: bool):
        monkeypatch.delenv("MYFM_HOST_RNG", raising=False)
        monkeypatch.delenv("MYFM_MT_FARM", raising=False)
        if host:
            monkeypatch.setenv("MYFM_HOST_RNG", "1")
        else:
            monkeypatch.setenv("MYFM_MT_FARM", "4,32768")
        with engine.engine_options(dtype=dtype):
            t = _TrainerHandle(X, [], y, 11, cfg)
            t.init_fm(32, 0.1)
        out = []
        for _ in range(7):
            t.step(1)
            t.sync()
            out.append(t.get_variates())
        return out

    dev, host = variates(False), variates(True)
    tol = 1e-6 if dtype == "f32" else 1e-14
    for it, (a, b) in enumerate(zip(dev, host)):
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=tol, atol=0, err_msg=f"sweep {it}")


def ordinal_1dim(n=1000):
    """The reference's ordered-probit fixture (tests/oprobit/test_oprobit_1dim.py:11-19)."""
    cps = np.asarray([0.0, 0.5, 1.5])
    rns = np.random.RandomState(0)
    X = rns.normal(0, 2, size=n)
    score = X * 0.5 + rns.randn(n)
    y = np.zeros(n)
    for cp in cps:
        y += (score > cp).astype(np.int64)
    return sps.csr_matrix(X[:, None]), y


def assert_ordered_close(trainer, chain, dtype, what, free_running=True):
    assert_state_close(trainer, chain, dtype, what, free_running=free_running)
    tol = FREE_RUNNING_F32 if dtype == "f32" else None
    close(trainer.get_fm()[3][0], chain.cutpoints()[0], dtype, f"cutpoints {what}", tol)
    assert trainer.mh_accept(0) == chain.mh_accept(0), what


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_ordered_probit_chain_1dim(engine, oracle, dtype):
    """Reference fixture, rank 0, 4 classes: cut-point Newton + MH on device-reduced sums, latent
    z from the shared mt19937 stream; compared with the oracle after init and every sweep."""
    X, y = ordinal_1dim()
    trainer, chain = make_pair(engine, oracle, X, y, 0, dtype, task="ordered", fit_w0=False)
    assert_ordered_close(trainer, chain, dtype, "after init")
    for it in range(8 if dtype == "f64" else 3):
        trainer.step(1)
        chain.step()
        assert_ordered_close(trainer, chain, dtype, f"after sweep {it}")


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_ordered_probit_chain_factorized(engine, oracle, dtype):
    """Ordered probit on MovieLens-shaped data (rank 4, 5 ordinal classes, two groups)."""
    X, y, group_shapes = movielens_like(6000, 150, 40, 3, seed=2)
    y = np.clip(np.round(y), 1, 5) - 1
    trainer, chain = make_pair(engine, oracle, X, y, 4, dtype, task="ordered", group_shapes=group_shapes)
    assert_ordered_close(trainer, chain, dtype, "after init")
    for it in range(6 if dtype == "f64" else 2):
        trainer.step(1)
        chain.step()
        assert_ordered_close(trainer, chain, dtype, f"after sweep {it}")


def test_ordered_probit_public_api(engine):
    """The reference's own test (tests/oprobit/test_oprobit_1dim.py:9-61) through MyFMOrderedProbit:
    planted cut-points recovered, predict_proba == manual Phi-differencing over the kept samples
    == running mean of the per-iteration callback."""
    from myfm_b200 import MyFMOrderedProbit
    from myfm_b200.base import std_cdf
    from myfm_b200.utils.callbacks import OrderedProbitCallback

    X, y = ordinal_1dim()
    with engine.engine_options(dtype="f64"):
        callback = OrderedProbitCallback(100, X_test=X, y_test=y, n_class=4)
        fm = MyFMOrderedProbit(0, fit_w0=False)
        fm.fit(X, y, callback=callback, n_iter=100, n_kept_samples=100)
        for cp_1, cp_2, cp_3 in fm.cutpoint_samples[-10:]:
            assert abs(cp_1) < 0.25 and abs(cp_2 - cp_1 - 0.5) < 0.25 and abs(cp_3 - cp_1 - 1.5) < 0.25
        p_core = fm.predict_proba(X)
        np.testing.assert_allclose(callback.predictions / 100, p_core, rtol=1e-7, atol=1e-12)
        manual = np.zeros((X.shape[0], 4))
        for sample in fm.predictor_.samples:
            score = sample.predict_score(X, [])
            cdf = std_cdf(sample.cutpoints[0][np.newaxis, :] - score[:, np.newaxis])
            diff = np.hstack([np.zeros((score.shape[0], 1)), cdf, np.ones((score.shape[0], 1))])
            manual += diff[:, 1:] - diff[:, :-1]
        manual /= len(fm.predictor_.samples)
        np.testing.assert_allclose(manual, p_core, rtol=1e-7, atol=1e-12)
        assert 0 < sum(fm.history_.n_mh_accept) <= 100 if hasattr(fm, "history_") else True
