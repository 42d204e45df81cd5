"""Host-side mirror of the reference extension surface (myfm_b200/_myfm.py): construction,
validation, error behaviour and pickling layouts — everything that needs no GPU."""
import pickle

import numpy as np
import pytest
import scipy.sparse as sps

from myfm_b200 import MyFMClassifier, MyFMOrderedProbit, MyFMRegressor, RelationBlock
from myfm_b200._myfm import (
    FM,
    ConfigBuilder,
    FMHyperParameters,
    LearningHistory,
    Predictor,
    TaskType,
    create_train_vfm,
    mean_var_truncated_normal_left,
    mean_var_truncated_normal_right,
)
from myfm_b200.base import check_data_consistency, std_cdf


def test_public_names_match_reference():
    import myfm_b200

    for name in ["RelationBlock", "MyFMOrderedProbit", "MyFMRegressor", "MyFMClassifier",
                 "MyFMGibbsRegressor", "MyFMGibbsClassifier"]:
        assert hasattr(myfm_b200, name)
    assert myfm_b200.MyFMRegressor is myfm_b200.MyFMGibbsRegressor
    from myfm_b200.utils.callbacks import (  # noqa: F401
        ClassificationCallback, LibFMLikeCallbackBase, OrderedProbitCallback, RegressionCallback)


def test_relation_block_surface_and_pickle():
    blk = RelationBlock(np.asarray([0, 2, 1, 1]), sps.csr_matrix(np.eye(3)))
    assert (blk.mapper_size, blk.block_size, blk.feature_size) == (4, 3, 3)
    assert blk.original_to_block == [0, 2, 1, 1]
    assert "mapper size = 4" in repr(blk) and "feature size = 3" in repr(blk)
    clone = pickle.loads(pickle.dumps([blk]))[0]
    assert clone.original_to_block == blk.original_to_block
    assert (clone.data != blk.data).nnz == 0
    with pytest.raises(RuntimeError, match="non-existing row"):
        RelationBlock([0, 3], sps.csr_matrix(np.eye(3)))
    with pytest.raises(TypeError):
        RelationBlock([0, -1], sps.csr_matrix(np.eye(3)))


def test_config_builder_validation():
    b = ConfigBuilder()
    assert b.set_alpha_0(2.0) is b and b.set_n_iter(3).set_n_kept_samples(2) is b
    assert b.set_identical_groups(5).build().n_groups == 1
    assert ConfigBuilder().set_group_index([0, 1, 1, 2]).build().n_groups == 3
    with pytest.raises(ValueError, match="No matching index for group index 1"):
        ConfigBuilder().set_group_index([0, 2]).build()
    with pytest.raises(ValueError, match="n_kept_samples must not exceed n_iter"):
        ConfigBuilder().set_identical_groups(2).set_n_iter(2).set_n_kept_samples(3).build()
    with pytest.raises(ValueError, match="n_iter must be positive"):
        ConfigBuilder().set_identical_groups(2).set_n_iter(0).set_n_kept_samples(0).build()
    with pytest.raises(ValueError, match="non-negative"):
        ConfigBuilder().set_identical_groups(2).set_n_kept_samples(-1).build()
    cfg = ConfigBuilder().set_identical_groups(2).set_task_type(TaskType.ORDERED).set_cutpoint_groups(
        [(3, [0, 1, 2])]).build()
    assert int(cfg.task_type) == 2


def test_state_holders_pickle_layouts():
    fm = FM(0.5, np.arange(3.0), np.arange(6.0).reshape(3, 2), [np.asarray([0.0, 1.0])])
    assert fm.__getstate__()[0] == 0.5 and len(fm.__getstate__()) == 4
    fm2 = pickle.loads(pickle.dumps(fm))
    np.testing.assert_array_equal(fm2.V, fm.V)
    np.testing.assert_array_equal(fm2.cutpoints[0], fm.cutpoints[0])
    legacy = FM.__new__(FM)
    legacy.__setstate__((0.1, np.zeros(3), np.zeros((3, 2))))  # 3-tuple of earlier versions
    assert legacy.cutpoints == []
    assert "feature size = 3, rank = 2" in repr(fm)

    hp = FMHyperParameters(1.5, np.zeros(2), np.ones(2), np.zeros((2, 3)), np.ones((2, 3)))
    hp2 = pickle.loads(pickle.dumps(hp))
    assert hp2.alpha == 1.5 and hp2.mu_V.shape == (2, 3)
    with pytest.raises(AttributeError):
        hp.alpha = 2.0  # def_readonly

    pred = Predictor(2, 3, int(TaskType.CLASSIFICATION))
    pred.samples.append(fm)
    state = pred.__getstate__()
    assert state[:3] == (2, 3, 1) and len(state[3]) == 1
    pred2 = pickle.loads(pickle.dumps(pred))
    assert pred2._type == TaskType.CLASSIFICATION and len(pred2.samples) == 1

    hist = LearningHistory()
    hist.hypers.append(hp)
    hist.n_mh_accept.append(3)
    hist2 = pickle.loads(pickle.dumps(hist))
    assert hist2.n_mh_accept == [3] and hist2.hypers[0].alpha == 1.5


def test_predictor_input_checks_need_no_gpu():
    pred = Predictor(2, 3, int(TaskType.REGRESSION))
    X = sps.csr_matrix(np.eye(4)[:, :3])
    with pytest.raises(RuntimeError, match="Empty samples!"):
        pred.predict(X, [])
    with pytest.raises(RuntimeError, match="no sample available"):
        pred.predict_parallel(X, [], 2)
    with pytest.raises(ValueError, match="Told to predict for 4 but this->feature_size is 3"):
        pred.predict(sps.csr_matrix(np.eye(4)), [])
    blk = RelationBlock([0, 0, 0], sps.csr_matrix(np.eye(2)))
    with pytest.raises(RuntimeError, match=r"main table has size 4 but the relation\[0\] has size 3"):
        pred.predict(X, [blk])


def test_estimators_before_fit_and_data_checks():
    m = MyFMRegressor(3)
    assert m.w0_samples is None and m.w_samples is None and m.V_samples is None
    with pytest.raises(RuntimeError, match="Predictor called before fit"):
        m.predict(sps.csr_matrix(np.eye(2)))
    with pytest.raises(RuntimeError, match="Sampler not run yet"):
        m.get_hyper_trace()
    assert "MyFMGibbsRegressor(init_stdev=0.1" in str(m)
    assert MyFMOrderedProbit(2).cutpoint_samples is None
    assert int(MyFMClassifier(2)._task_type) == 1
    with pytest.raises(ValueError, match="At least X or X_rel"):
        check_data_consistency(None, [])
    a = RelationBlock([0, 1], sps.csr_matrix(np.eye(2)))
    b = RelationBlock([0, 1, 1], sps.csr_matrix(np.eye(2)))
    with pytest.raises(ValueError, match="Inconsistent case size"):
        check_data_consistency(None, [a, b])
    with pytest.raises(ValueError, match="different shape"):
        check_data_consistency(sps.csr_matrix(np.eye(3)), [a])
    assert check_data_consistency(None, [a]) == 2
    with pytest.raises(RuntimeError, match="Must specify both"):
        m.fit(sps.csr_matrix(np.eye(2)), np.zeros(2), y_test=np.zeros(2), n_iter=1)


def test_truncated_normal_moments():
    from scipy import stats

    for mu in (-3.0, -0.5, 0.0, 0.7, 4.0):
        mean, var, lnZ = mean_var_truncated_normal_left(mu)
        tn = stats.truncnorm(-mu, np.inf, loc=mu)
        assert abs(mean - tn.mean()) < 1e-8 and abs(var - tn.var()) < 1e-8
        assert abs(lnZ - (stats.norm.logcdf(mu) + np.log(2))) < 1e-8
        rmean, rvar, _ = mean_var_truncated_normal_right(mu)
        tn = stats.truncnorm(-np.inf, -mu, loc=mu)
        assert abs(rmean - tn.mean()) < 1e-8 and abs(rvar - tn.var()) < 1e-8
    assert abs(std_cdf(np.asarray([0.0]))[0] - 0.5) < 1e-15
    with pytest.raises(NotImplementedError):
        create_train_vfm()


def test_engine_options():
    import myfm_b200

    before = myfm_b200.get_options()
    with myfm_b200.engine_options(dtype="f32") as o:
        assert o.dtype == "f32" and myfm_b200.get_options().dtype == "f32"
    assert myfm_b200.get_options() == before
    with pytest.raises(ValueError):
        myfm_b200.set_options(dtype="f16")


def test_import_myfm_drop_in():
    """The reference's package name resolves to the engine's modules of the same names."""
    import myfm
    import myfm_b200
    from myfm import MyFMClassifier, MyFMOrderedProbit, MyFMRegressor, RelationBlock  # noqa: F401
    from myfm._myfm import ConfigBuilder, create_train_fm  # noqa: F401
    from myfm.utils.callbacks import ClassificationCallback, OrderedProbitCallback, RegressionCallback  # noqa: F401
    import myfm.gibbs

    assert MyFMRegressor is myfm_b200.MyFMRegressor and myfm.gibbs is myfm_b200.gibbs
    assert set(myfm.__all__) >= {"RelationBlock", "MyFMOrderedProbit", "MyFMRegressor", "MyFMClassifier",
                                 "MyFMGibbsRegressor", "MyFMGibbsClassifier"}


def test_create_train_fm_loop_order_with_observer_callbacks(monkeypatch):
    """The per-iteration protocol of create_train_fm (FMTrainer.hpp:56-87) on a recording stand-in for the
    device trainer: step -> kept-sample copy -> hyper fetch -> callback; an `observer` callback has the next
    sweep queued before it runs, a stop request ends the loop, and no sweep is ever observed twice."""
    from myfm_b200 import _myfm

    log = []

    class FakeTrainer:
        def __init__(self, X, relations, y, seed, config):
            self.dim_all, self.n_cutpoint_groups, self.steps = 3, 0, 0

        def init_fm(self, rank, init_std):
            log.append("init")

        def step(self, n):
            self.steps += n
            log.append(f"step{self.steps}")

        def snapshot(self):
            log.append(f"snap{self.steps}")
            return self.steps

        def get_hyper(self):
            log.append(f"hyper{self.steps}")
            return self.steps

        def sync(self):
            log.append("sync")

    monkeypatch.setattr(_myfm, "_TrainerHandle", FakeTrainer)
    monkeypatch.delenv("MYFM_B200_NO_RUN_AHEAD", raising=False)
    monkeypatch.delenv("MYFM_NO_RUN_AHEAD", raising=False)
    config = ConfigBuilder().set_identical_groups(3).set_n_iter(4).set_n_kept_samples(2).build()

    def run(observer, stop_at=None):
        log.clear()
        seen = []

        def callback(i, fm, hyper, history):
            log.append(f"cb{i}")
            seen.append((i, hyper, len(history.hypers)))
            return stop_at == i

        if observer is not None:
            callback.observer = observer
        predictor, history = _myfm.create_train_fm(2, 0.1, None, [], None, 0, config, callback)
        return list(log), seen, predictor.samples, history.hypers

    plain = run(None)
    assert plain[0] == ["init", "step1", "hyper1", "cb0", "step2", "hyper2", "cb1", "step3", "snap3", "hyper3", "cb2",
                        "step4", "snap4", "hyper4", "cb3", "sync"]
    ahead = run(True)
    assert ahead[0] == ["init", "step1", "hyper1", "step2", "cb0", "hyper2", "step3", "cb1", "snap3", "hyper3",
                        "step4", "cb2", "snap4", "hyper4", "cb3", "sync"]
    assert plain[1:] == ahead[1:] == run(lambda i: i % 2 == 0)[1:]  # same callbacks, kept samples, history
    assert plain[1] == [(0, 1, 1), (1, 2, 2), (2, 3, 3), (3, 4, 4)] and plain[2] == [3, 4]
    # a stop request: same callbacks and samples; the observer variant ran one unobserved sweep ahead
    stop_plain, stop_ahead = run(None, stop_at=2), run(True, stop_at=2)
    assert stop_plain[1:] == stop_ahead[1:] and stop_plain[2] == [3]
    assert stop_plain[0][-3:] == ["hyper3", "cb2", "sync"] and stop_ahead[0][-4:] == ["hyper3", "step4", "cb2", "sync"]
    monkeypatch.setenv("MYFM_B200_NO_RUN_AHEAD", "1")
    assert run(True)[0] == plain[0]
