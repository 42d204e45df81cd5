"""LibFM-style per-iteration callbacks (reference src/myfm/utils/callbacks/libfm.py:57-262) on the device
(csrc/eval_device.cuh through myfm_evaluator_*): every metric of every sweep equals the numpy path of the
reference fed with the same sample, the running sums equal numpy's, and the reference's identity
`predict(X_test) == callback.predictions / n_iter` holds (tests/regression/test_fit.py:39)."""
import numpy as np
import pytest

from helpers import movielens_like

pytestmark = pytest.mark.gpu


def _both(device_cb, host_cb):
    def callback(i, fm, hyper, history):
        out = device_cb(i, fm, hyper, history)
        host_cb(i, fm.freeze(), hyper, history)  # a plain host FM: the reference's numpy path
        return out

    return callback


def _compare(device_cb, host_cb, n_iter):
    assert device_cb._device is not None and host_cb._device is None
    assert len(device_cb.result_trace) == len(host_cb.result_trace) == n_iter
    for it, (a, b) in enumerate(zip(device_cb.result_trace, host_cb.result_trace)):
        assert list(a) == list(b)
        for key in a:
            if np.isnan(b[key]):
                assert np.isnan(a[key]), (it, key)
            else:
                np.testing.assert_allclose(a[key], b[key], rtol=1e-9, atol=1e-12, err_msg=f"sweep {it} {key}")
    np.testing.assert_allclose(device_cb.predictions, host_cb.predictions, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(device_cb.prediction_all_but_5, host_cb.prediction_all_but_5, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_regression_callback_on_device(engine, dtype):
    from myfm_b200.utils.callbacks import RegressionCallback

    X, y, gs = movielens_like(6000, 80, 40, 3, seed=31)
    Xtr, ytr, Xte, yte = X[:5000], y[:5000], X[5000:], y[5000:]
    n_iter = 12
    dev = RegressionCallback(n_iter, Xte, yte, clip_min=1.0, clip_max=5.0)
    host = RegressionCallback(n_iter, Xte, yte, clip_min=1.0, clip_max=5.0)
    with engine.engine_options(dtype=dtype):
        fm = engine.MyFMRegressor(rank=4, random_seed=3).fit(Xtr, ytr, n_iter=n_iter, n_kept_samples=n_iter,
                                                             group_shapes=gs, callback=_both(dev, host))
        _compare(dev, host, n_iter)
        # (the f32 predictor accumulates the kept samples in f32, the callback in f64)
        np.testing.assert_allclose(fm.predict(Xte), dev.predictions / n_iter, rtol=1e-7 if dtype == "f64" else 2e-6)
    assert dev.result_trace[-1]["rmse"] < 1.3


def test_classification_callback_on_device(engine):
    from myfm_b200.utils.callbacks import ClassificationCallback

    X, y, gs = movielens_like(6000, 80, 40, 3, seed=32)
    yb = (y > np.median(y)).astype(np.float64)
    Xtr, ytr, Xte, yte = X[:5000], yb[:5000], X[5000:], yb[5000:]
    n_iter = 10
    dev, host = ClassificationCallback(n_iter, Xte, yte), ClassificationCallback(n_iter, Xte, yte)
    with engine.engine_options(dtype="f64"):
        fm = engine.MyFMClassifier(rank=3, random_seed=4).fit(Xtr, ytr, n_iter=n_iter, n_kept_samples=n_iter,
                                                              group_shapes=gs, callback=_both(dev, host))
        _compare(dev, host, n_iter)
        np.testing.assert_allclose(fm.predict_proba(Xte), dev.predictions / n_iter, rtol=1e-7)


def test_ordered_probit_callback_on_device(engine):
    from myfm_b200.utils.callbacks import OrderedProbitCallback

    X, y, gs = movielens_like(6000, 80, 40, 3, seed=33)
    yo = np.digitize(y, np.quantile(y, [0.25, 0.5, 0.75])).astype(np.float64)
    Xtr, ytr, Xte, yte = X[:5000], yo[:5000], X[5000:], yo[5000:]
    n_iter = 10
    dev, host = OrderedProbitCallback(n_iter, Xte, yte, 4), OrderedProbitCallback(n_iter, Xte, yte, 4)
    with engine.engine_options(dtype="f64"):
        fm = engine.MyFMOrderedProbit(rank=3, random_seed=5).fit(Xtr, ytr, n_iter=n_iter, n_kept_samples=n_iter,
                                                                 group_shapes=gs, callback=_both(dev, host))
        _compare(dev, host, n_iter)
        np.testing.assert_allclose(fm.predict_proba(Xte), dev.predictions / n_iter, rtol=1e-7, atol=1e-12)


def test_observer_callback_runs_the_same_chain(engine):
    """create_train_fm starts the next sweep before an `observer` callback runs (myfm_b200/_myfm.py): the chain,
    its history, the kept samples and an early stop are those of the plain loop."""
    X, y, gs = movielens_like(6000, 80, 40, 3, seed=32)

    def fit(observer, stop_at=None):
        seen = []

        def callback(i, fm, hyper, history):
            seen.append((i, hyper.alpha, len(history.hypers)))
            return (stop_at is not None and i == stop_at), None

        callback.observer = observer
        with engine.engine_options(dtype="f64"):
            model = engine.MyFMRegressor(rank=4, random_seed=5).fit(X, y, n_iter=12, n_kept_samples=5,
                                                                    group_shapes=gs, callback=callback)
        return model, seen

    plain, seen_plain = fit(False)
    ahead, seen_ahead = fit(True)
    odd, seen_odd = fit(lambda i: i % 2 == 1)
    assert seen_plain == seen_ahead == seen_odd and len(seen_plain) == 12
    for other in (ahead, odd):
        np.testing.assert_array_equal(plain.predict(X[:500]), other.predict(X[:500]))
        assert [h.alpha for h in plain.history_.hypers] == [h.alpha for h in other.history_.hypers]
    # early stop: the callbacks seen and the samples kept are the same (one sweep ran ahead, unobserved)
    stop_plain, s1 = fit(False, stop_at=9)
    stop_ahead, s2 = fit(True, stop_at=9)
    assert s1 == s2 and len(s1) == 10
    np.testing.assert_array_equal(stop_plain.predict(X[:500]), stop_ahead.predict(X[:500]))
    # fit()'s own progress callback is an observer between its status lines
    with engine.engine_options(dtype="f64"):
        default = engine.MyFMRegressor(rank=4, random_seed=5).fit(X, y, n_iter=12, n_kept_samples=5, group_shapes=gs)
    np.testing.assert_array_equal(plain.predict(X[:500]), default.predict(X[:500]))
