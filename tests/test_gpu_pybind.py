"""The compiled pybind11 binding over the C ABI (myfm_b200/csrc/pybind_binding.cpp -> myfm_b200._myfm_pybind):
the reference's `create_train_fm` entry point (cpp_source/declare_module.hpp:30-45) with array_t arguments.
It must walk the same chain as the ctypes binding and hand back the same types."""
import numpy as np
import pytest

from helpers import block_data, movielens_like

pytestmark = pytest.mark.gpu


def _config(group_shapes, n_iter, task=None, y=None):
    from myfm_b200._myfm import ConfigBuilder, TaskType

    b = (ConfigBuilder().set_mu_0(0.0).set_group_index(np.repeat(np.arange(len(group_shapes)), group_shapes))
         .set_n_iter(n_iter).set_n_kept_samples(n_iter))
    if task == "ordered":
        b.set_task_type(TaskType.ORDERED).set_cutpoint_groups([(int(y.max()) + 1, np.arange(y.shape[0]))])
    return b.build()


def test_pybind_create_train_fm_equals_ctypes(engine):
    from myfm_b200 import _myfm, _myfm_pybind
    from myfm_b200.csrc import build

    build.build_pybind()
    X, y, gs = movielens_like(5000, 70, 30, 3, seed=51)
    n_iter = 6
    seen = []
    with engine.engine_options(dtype="f64"):
        pa, ha = _myfm_pybind.create_train_fm(4, 0.1, X, [], y, 7, _config(gs, n_iter),
                                              lambda i, fm, hyper, hist: seen.append((i, fm.w0, hyper.alpha)) or False)
        pb, hb = _myfm.create_train_fm(4, 0.1, X, [], y, 7, _config(gs, n_iter), lambda i, fm, hyper, hist: False)
        assert type(pa) is _myfm.Predictor and type(ha) is _myfm.LearningHistory
        assert len(pa.samples) == len(pb.samples) == n_iter and len(seen) == n_iter
        for sa, sb in zip(pa.samples, pb.samples):
            assert sa.w0 == sb.w0
            np.testing.assert_array_equal(sa.w, sb.w)
            np.testing.assert_array_equal(sa.V, sb.V)
        for xa, xb in zip(ha.hypers, hb.hypers):
            assert xa.alpha == xb.alpha
            np.testing.assert_array_equal(xa.mu_V, xb.mu_V)
        np.testing.assert_allclose(pa.predict(X[:100], []), pb.predict(X[:100], []), rtol=1e-12)
        s = pa.samples[-1]
        np.testing.assert_allclose(_myfm_pybind.predict_score(s.w0, s.w, s.V, X[:100], []),
                                   s.predict_score(X[:100], []), rtol=1e-12)


def test_pybind_relation_blocks_and_ordered_probit(engine):
    from myfm_b200 import _myfm, _myfm_pybind

    X_flat, tm, (ui, ub), (ii, ib), y, gs = block_data(300)
    yo = np.digitize(y, np.quantile(y, [0.3, 0.7])).astype(np.float64)
    rels = [_myfm.RelationBlock(ui, ub), _myfm.RelationBlock(ii, ib)]
    with engine.engine_options(dtype="f64"):
        pa, ha = _myfm_pybind.create_train_fm(2, 0.1, tm, rels, yo, 3, _config(gs, 5, "ordered", yo),
                                              lambda *a: False)
        pb, hb = _myfm.create_train_fm(2, 0.1, tm, rels, yo, 3, _config(gs, 5, "ordered", yo), lambda *a: False)
    assert ha.n_mh_accept == hb.n_mh_accept
    for sa, sb in zip(pa.samples, pb.samples):
        np.testing.assert_array_equal(sa.V, sb.V)
        np.testing.assert_array_equal(sa.cutpoints[0], sb.cutpoints[0])


def test_pybind_errors_map_like_the_reference(engine):
    from myfm_b200 import _myfm_pybind

    X, y, gs = movielens_like(200, 10, 5, 2, seed=52)
    with pytest.raises(RuntimeError, match="Shape mismatch"):  # BaseFMTrainer.hpp:69-76
        _myfm_pybind.create_train_fm(2, 0.1, X, [], y[:-1], 1, _config(gs, 2), lambda *a: False)
