"""Kept samples resident on the device (myfm_trainer_snapshot / myfm_predict_samples_mean): the
predictor over them equals the reference's definition (mean over the kept samples of the
per-sample score, predictor.hpp:126-147) computed from their host copies, survives pickling, and
falls back to the host path when the engine options change."""
import pickle

import numpy as np
import pytest

from helpers import block_data, fm_prediction, FMWeights, movielens_like

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_regressor_predict_over_device_samples(engine, dtype):
    from myfm_b200 import MyFMRegressor
    from myfm_b200._myfm import FM, _DeviceFM

    X, y, gs = movielens_like(6000, 120, 40, 3, seed=2)
    with engine.engine_options(dtype=dtype):
        fm = MyFMRegressor(rank=4, random_seed=1).fit(X, y, n_iter=12, n_kept_samples=5, group_shapes=gs)
        samples = fm.predictor_.samples
        assert len(samples) == 5 and all(isinstance(s, _DeviceFM) for s in samples)
        on_device = fm.predict(X[:2000])
        # per-sample scores from the host copies, plain numpy
        manual = np.mean([fm_prediction(X[:2000], FMWeights(s.w0, s.w, s.V.T)) for s in samples], axis=0)
        tol = 1e-9 if dtype == "f64" else 2e-5
        np.testing.assert_allclose(on_device, manual, rtol=tol, atol=tol)
        # a single kept sample predicts on the device too
        one = samples[0].predict_score(X[:500], [])
        np.testing.assert_allclose(one, fm_prediction(X[:500], FMWeights(samples[0].w0, samples[0].w, samples[0].V.T)),
                                   rtol=tol, atol=tol)
        # pickling turns the samples into plain FMs; the host path gives the same mean
        clone = pickle.loads(pickle.dumps(fm))
        assert all(type(s) is FM for s in clone.predictor_.samples)
        np.testing.assert_allclose(clone.predict(X[:2000]), on_device, rtol=tol, atol=tol)
    # other engine options (dtype of the prediction dataset differs): host path, same numbers
    other = "f32" if dtype == "f64" else "f64"
    with engine.engine_options(dtype=other):
        np.testing.assert_allclose(fm.predict(X[:2000]), on_device, rtol=2e-5, atol=2e-5)


def test_kept_samples_equal_callback_states(engine):
    """The device-to-device copy is the sample of that very iteration (FMTrainer.hpp:74-76)."""
    from myfm_b200 import MyFMRegressor

    X, y, gs = movielens_like(3000, 60, 30, 2, seed=4)
    seen = []

    def callback(i, fm, hyper, history):
        seen.append((fm.w0, fm.w.copy(), fm.V.copy()))
        return False, None

    with engine.engine_options(dtype="f64"):
        fm = MyFMRegressor(rank=3, random_seed=3).fit(X, y, n_iter=8, n_kept_samples=3, group_shapes=gs,
                                                      callback=callback)
    for (w0, w, V), s in zip(seen[-3:], fm.predictor_.samples):
        assert w0 == s.w0
        np.testing.assert_array_equal(w, s.w)
        np.testing.assert_array_equal(V, s.V)


def test_relation_blocks_and_classifier(engine):
    from myfm_b200 import MyFMClassifier
    from myfm_b200._myfm import RelationBlock
    from myfm_b200.base import std_cdf

    X_flat, main, users, items, y, gs = block_data(300)
    blocks = [RelationBlock(*users), RelationBlock(*items)]
    yb = y > np.median(y)
    with engine.engine_options(dtype="f64"):
        fm = MyFMClassifier(rank=2, random_seed=5).fit(main, yb, X_rel=blocks, n_iter=10, n_kept_samples=4,
                                                       group_shapes=gs)
        p = fm.predict_proba(main, X_rel=blocks)
        manual = np.mean([std_cdf(fm_prediction(X_flat, FMWeights(s.w0, s.w, s.V.T))) for s in fm.predictor_.samples],
                         axis=0)
    np.testing.assert_allclose(p, manual, rtol=1e-9, atol=1e-9)
