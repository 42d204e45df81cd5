"""Pins the oracle (oracle/fm_oracle.hpp) with what the reference's own test-suite pins for this
path, re-run against the restatement (the reference itself cannot be built here: no Eigen):

* tests/regression/test_block.py:80-149   relation blocks == flattened design, rtol 1e-7
* tests/regression/test_fit.py:20-72      planted-parameter recovery (alpha, w0, w, V cross-terms)
* tests/classification/test_classification.py:14-70   cross-term recovery under the probit link
* tests/oprobit/test_oprobit_1dim.py:9-61 cut-points within 0.25 of the planted ones
* libstdc++ <random> known answers (the RNG is part of the results contract)
"""
import numpy as np
import pytest
import scipy.sparse as sps

from helpers import STUB_WEIGHT, block_data, dense_block_data, fm_prediction, middle_data, toy_matrix


def test_libstdcxx_known_answers(oracle):
    # mt19937(42): the first two draws of one persistent normal_distribution
    np.testing.assert_allclose(oracle.kat_normal("f64", 42, 2),
                               [-0.55023449442049355, 0.51543306969120128], rtol=1e-15)
    np.testing.assert_allclose(oracle.kat_normal("f32", 42, 2), [1.22192132, -0.516964138], rtol=1e-7)


@pytest.mark.parametrize("alpha_inv", [0.3, 1.0, 3])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_planted_regression(oracle, alpha_inv, dtype):
    X, score = middle_data()
    y = score + alpha_inv * np.random.RandomState(0).normal(0, 1, size=score.shape)
    chain = oracle.OracleChain(X, y, 3, dtype=dtype, n_iter=100, n_kept_samples=100)
    samples, hypers = chain.run()
    assert len(samples) == 100 and len(hypers) == 100
    last_alphas = np.asarray([h["alpha"] for h in hypers[-20:]])
    assert np.all(last_alphas > (1 / alpha_inv ** 2) / 2) and np.all(last_alphas < (1 / alpha_inv ** 2) * 2)
    for w0, w, V, _ in samples[-20:]:
        assert abs(w0 - STUB_WEIGHT.global_bias) < 0.5
        assert np.all(np.abs(w - STUB_WEIGHT.weight) < 1.0)
        for i in range(3):
            for j in range(i + 1, 3):
                cross = STUB_WEIGHT.factors[:, i].dot(STUB_WEIGHT.factors[:, j])
                if abs(cross) < 0.1:
                    continue
                sign = cross / abs(cross)
                assert sign * cross * 0.5 < V[i].dot(V[j]) < sign * cross * 2
    # the posterior mean of the forward pass reproduces the data up to the noise level
    pred = np.mean([oracle.predict_score(dtype, s[0], s[1], s[2], X) for s in samples[5:]], axis=0)
    assert np.sqrt(np.mean((pred - y) ** 2)) < 1.1 * alpha_inv


def test_predict_score_matches_numpy(oracle):
    X, _ = middle_data(200)
    got = oracle.predict_score("f64", STUB_WEIGHT.global_bias, STUB_WEIGHT.weight, STUB_WEIGHT.factors.T, X)
    np.testing.assert_allclose(got, fm_prediction(X, STUB_WEIGHT), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("maker", [block_data, dense_block_data])
def test_block_equals_flat(oracle, maker):
    X_flat, main, (u_idx, u_blk), (i_idx, i_blk), y, group_shapes = maker()
    kw = dict(dtype="f64", fit_w0=False, group_shapes=group_shapes, n_iter=30, n_kept_samples=30)
    flat, _ = oracle.OracleChain(X_flat, y, 2, **kw).run()
    blocked, _ = oracle.OracleChain(main, y, 2, X_rel=[(u_idx, u_blk), (i_idx, i_blk)], **kw).run()
    for s_flat, s_blk in zip(flat, blocked):
        np.testing.assert_allclose(s_flat[2], s_blk[2], rtol=1e-7)
        np.testing.assert_allclose(s_flat[1], s_blk[1], rtol=1e-7)
    # forward pass: blocks == flat
    w0, w, V, _ = flat[-1]
    np.testing.assert_allclose(
        oracle.predict_score("f64", w0, w, V, X_flat),
        oracle.predict_score("f64", w0, w, V, main, [(u_idx, u_blk), (i_idx, i_blk)]), rtol=1e-9)


def test_planted_classification(oracle):
    X, score = middle_data()
    rns = np.random.RandomState(0)
    noised = score + rns.normal(0, 1, size=score.shape)
    noised -= noised.mean()
    y = (noised > 0).astype(np.float64) * 2 - 1  # ClassifierMixin._process_y
    chain = oracle.OracleChain(X, y, 3, dtype="f64", task="classification", n_iter=200, n_kept_samples=200)
    samples, hypers = chain.run()
    assert all(h["alpha"] == 1.0 for h in hypers)
    for _, _, V, _ in samples[-20:]:
        for i in range(3):
            for j in range(i + 1, 3):
                cross = STUB_WEIGHT.factors[:, i].dot(STUB_WEIGHT.factors[:, j])
                if abs(cross) < 0.5:
                    continue
                sign = cross / abs(cross)
                assert sign * cross * 0.5 < V[i].dot(V[j]) < sign * cross * 2


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_ordered_probit_cutpoints(oracle, dtype):
    n = 1000
    cps = np.asarray([0.0, 0.5, 1.5])
    rns = np.random.RandomState(0)
    X = rns.normal(0, 2, size=n)
    score = X * 0.5 + rns.randn(n)
    y = np.zeros(n)
    for cp in cps:
        y += (score > cp).astype(np.int64)
    chain = oracle.OracleChain(sps.csr_matrix(X[:, None]), y, 0, dtype=dtype, task="ordered",
                               fit_w0=False, n_iter=100, n_kept_samples=100)
    samples, _ = chain.run()
    for _, _, _, cutpoints in samples[-10:]:
        c1, c2, c3 = cutpoints[0]
        assert abs(c1) < 0.25 and abs(c2 - c1 - 0.5) < 0.25 and abs(c3 - c1 - 1.5) < 0.25
    assert 0 < chain.mh_accept(0) <= 100


def test_error_behaviour(oracle):
    X, y = toy_matrix()
    with pytest.raises(ValueError, match="No matching index for group index 1"):
        oracle.OracleChain(X, y, 2, group_index=[0, 2] + [0] * 7)
    with pytest.raises(RuntimeError, match="Shape mismatch"):
        oracle.OracleChain(X, y[:3], 2)
    with pytest.raises(RuntimeError, match="non-existing row"):
        oracle.OracleChain(X, y, 2, X_rel=[(np.asarray([0, 1, 2, 3]), sps.eye(3).tocsr())],
                           group_index=[0] * 12)
    with pytest.raises(ValueError, match="n_kept_samples must not exceed n_iter"):
        oracle.OracleChain(X, y, 2, n_iter=3, n_kept_samples=4)


def test_fit_w0_false_keeps_stale_bias_until_update_e(oracle):
    """FMTrainer.hpp:219-222: w0 is zeroed but e is only corrected by the closing update_e."""
    X, y = toy_matrix()
    chain = oracle.OracleChain(X, y, 2, fit_w0=False, n_iter=3)
    chain.step()
    w0, w, V = chain.fm()
    assert w0 == 0.0
    np.testing.assert_allclose(chain.e(), oracle.predict_score("f64", 0.0, w, V, X) - y, rtol=1e-12, atol=1e-12)


def test_truncated_normal_support(oracle):
    for dtype in ("f64", "f32"):
        assert np.all(oracle.tn_draws(dtype, 1, 0, -0.7, 0, 500) > -0.7)
        assert np.all(oracle.tn_draws(dtype, 1, 0, 2.5, 0, 500) > 2.5)
        assert np.all(oracle.tn_draws(dtype, 1, 1, 0.3, 0, 500) < 0.3)
        z = oracle.tn_draws(dtype, 1, 2, 0.2, 0.9, 500)
        assert np.all((z >= 0.2) & (z <= 0.9))
    # mean of N(0,1) truncated to (1, inf) is phi(1)/(1-Phi(1)) = 1.5251...
    assert abs(oracle.tn_draws("f64", 7, 0, 1.0, 0, 20000).mean() - 1.52513528) < 0.02
