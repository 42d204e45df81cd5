"""rng="philox": the per-row latent draws of classification / ordered probit run on the device from
counter-based streams (csrc/latent_device.cuh) instead of row by row on the host.  The chain is
statistically equivalent to the reference's, not seed-identical, so these tests are the
reference's own statistical tests (planted-parameter recovery, cut-points within 0.25:
tests/classification/test_classification.py:14-70, tests/oprobit/test_oprobit_1dim.py:9-61) plus
the exact distribution of the draws: support and mean of every truncated normal against the
closed form."""
import numpy as np
import pytest
import scipy.sparse as sps
from scipy import stats

from helpers import STUB_WEIGHT, fm_prediction, FMWeights, middle_data, movielens_like
from test_gpu_parity import make_pair, ordinal_1dim

pytestmark = pytest.mark.gpu


def _handle(engine, X, y, rank, task, dtype="f64", gs=None, n_iter=10, **kw):
    from myfm_b200._myfm import ConfigBuilder, TaskType, _TrainerHandle

    b = ConfigBuilder().set_mu_0(0.0).set_n_iter(n_iter).set_n_kept_samples(n_iter)
    b.set_fit_w0(kw.get("fit_w0", True))
    if gs is None:
        b.set_identical_groups(X.shape[1])
    else:
        b.set_group_index(np.repeat(np.arange(len(gs)), gs))
    b.set_task_type({"classification": TaskType.CLASSIFICATION, "ordered": TaskType.ORDERED,
                     "regression": TaskType.REGRESSION}[task])
    if task == "ordered":
        b.set_cutpoint_groups([(int(y.max()) + 1, np.arange(y.shape[0]))])
    with engine.engine_options(dtype=dtype, rng="philox"):
        t = _TrainerHandle(X, [], y, 42, b.build())
        t.init_fm(rank, 0.1)
    return t


def _score(t, X):
    w0, w, V, _ = t.get_fm()
    return fm_prediction(X, FMWeights(w0, w, V.T))


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_classification_latents_follow_the_truncated_normal(engine, dtype):
    X, y, gs = movielens_like(200_000, 300, 100, 2, seed=1)
    y = np.where(y > np.median(y), 1.0, -1.0)
    t = _handle(engine, X, y, 2, "classification", dtype, gs)
    t.step(1)
    pred = _score(t, X)
    z = pred - t.get_e()  # e = score - z (FMTrainer.hpp:498-512)
    assert np.all(z[y > 0] > 0) and np.all(z[y < 0] < 0)
    # E[z] of N(pred, 1) truncated at 0: pred +- phi(pred) / Phi(+-pred)
    expect = np.where(y > 0, pred + stats.norm.pdf(pred) / stats.norm.cdf(pred),
                      pred - stats.norm.pdf(pred) / stats.norm.cdf(-pred))
    assert abs(np.mean(z - expect)) < 0.01
    var = np.where(y > 0, stats.truncnorm.var(-pred, np.inf), stats.truncnorm.var(-np.inf, -pred))
    assert abs(np.mean((z - expect) ** 2) / np.mean(var) - 1) < 0.03
    # far tails use the exponential proposal: still inside the support, right mean
    tail = (np.abs(pred) > 1.0) & (np.sign(pred) != y)
    if tail.sum() > 500:
        assert abs(np.mean((z - expect)[tail])) < 0.05


def test_ordered_latents_follow_the_truncated_normal(engine):
    rng = np.random.default_rng(0)
    n = 100_000
    x = rng.normal(0, 2, n)
    y = np.digitize(0.5 * x + rng.normal(size=n), [0.0, 0.5, 1.5]).astype(np.float64)
    X = sps.csr_matrix(x[:, None])
    t = _handle(engine, X, y, 0, "ordered", "f64", fit_w0=False)
    t.step(1)
    gamma = t.get_fm()[3][0]
    pred = _score(t, X)
    z = pred - t.get_e()
    lo = np.concatenate([[-np.inf], gamma])[y.astype(int)]
    hi = np.concatenate([gamma, [np.inf]])[y.astype(int)]
    assert np.all((z > lo) & (z < hi))
    expect = stats.truncnorm.mean(lo - pred, hi - pred, loc=pred)
    assert abs(np.mean(z - expect)) < 0.01


def test_classifier_recovers_planted_factors(engine):
    """tests/classification/test_classification.py:14-70 with the device latent draws."""
    from myfm_b200 import MyFMClassifier

    X, score = middle_data()
    noised = score + np.random.RandomState(0).normal(0, 1, size=score.shape)
    noised -= noised.mean()
    with engine.engine_options(dtype="f64", rng="philox"):
        fm = MyFMClassifier(3).fit(X, noised > 0, n_iter=200, n_kept_samples=50)
        p = fm.predict_proba(X)
    assert np.mean((p > 0.5) == (noised > 0)) > 0.75
    for s in fm.predictor_.samples[-20:]:
        V = s.V
        for i in range(3):
            for j in range(i + 1, 3):
                cross = STUB_WEIGHT.factors[:, i].dot(STUB_WEIGHT.factors[:, j])
                if abs(cross) < 0.5:
                    continue
                sign = cross / abs(cross)
                assert sign * cross * 0.5 < V[i].dot(V[j]) < sign * cross * 2


def test_ordered_probit_recovers_cutpoints(engine):
    from myfm_b200 import MyFMOrderedProbit

    X, y = ordinal_1dim()
    with engine.engine_options(dtype="f64", rng="philox"):
        fm = MyFMOrderedProbit(0, fit_w0=False).fit(X, y, n_iter=100, n_kept_samples=100)
        for c1, c2, c3 in fm.cutpoint_samples[-10:]:
            assert abs(c1) < 0.25 and abs(c2 - c1 - 0.5) < 0.25 and abs(c3 - c1 - 1.5) < 0.25
        assert fm.predict_proba(X).shape == (X.shape[0], 4)


def test_regression_is_the_same_chain_in_both_modes(engine, oracle):
    """Regression has no latent draws: rng="philox" leaves the mt19937 chain untouched."""
    X, y, gs = movielens_like(5000, 80, 30, 3, seed=2)
    a = _handle(engine, X, y, 4, "regression", "f64", gs)
    b, _ = make_pair(engine, oracle, X, y, 4, "f64", group_shapes=gs)
    a.step(3), b.step(3)
    for xa, xb in zip(a.get_fm()[:3], b.get_fm()[:3]):
        np.testing.assert_array_equal(xa, xb)
