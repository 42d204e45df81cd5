"""Gibbs-sampled Bayesian Factorization Machines: regressor, probit classifier, ordered probit.
Public surface of the reference's ``src/myfm/gibbs.py`` over the CUDA engine.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse as sps

from ._myfm import (
    FM,
    ConfigBuilder,
    FMHyperParameters,
    FMLearningConfig,
    LearningHistory,
    Predictor,
    RelationBlock,
    TaskType,
    create_train_fm,
)
from .base import (
    REAL,
    ArrayLike,
    ClassifierMixin,
    DenseArray,
    MyFMBase,
    RegressorMixin,
    _main_table,
    check_data_consistency,
)

_FIT_DOC = """Run the Gibbs sampler on (X, X_rel, y).

        X: 2D array-like main table (or None when only relation blocks are given).
        y: 1D targets.  X_rel: relation blocks supplementing X.
        X_test, y_test, X_rel_test: optional held-out data reported by the default callback.
        n_iter: number of sweeps.  n_kept_samples: samples kept for prediction
        (default n_iter - 5, at least 5, at most n_iter).
        grouping: group id of every column (w_i, V_ir of group g share the hyper-priors
        mu_w[g], lambda_w[g], mu_V[g, r], lambda_V[g, r]); group_shapes: sizes of consecutive
        groups, ignored when grouping is given.
        callback(i, fm, hyper, history) -> (stop, message): called after every sweep.
        """


class MyFMGibbsBase(MyFMBase[FM, FMHyperParameters, Predictor, LearningHistory]):
    def _stack_samples(self, getter) -> Optional[DenseArray]:
        if self.predictor_ is None:
            return None
        return np.asarray([getter(fm) for fm in self.predictor_.samples], dtype=np.float64)

    @property
    def w0_samples(self) -> Optional[DenseArray]:
        """Kept samples of the global bias, or None before fit (reference gibbs.py:40-49)."""
        return self._stack_samples(lambda fm: fm.w0)

    @property
    def w_samples(self) -> Optional[DenseArray]:
        """Kept samples of the linear coefficients, (sample, feature) (reference gibbs.py:51-61)."""
        return self._stack_samples(lambda fm: fm.w)

    @property
    def V_samples(self) -> Optional[DenseArray]:
        """Kept samples of the factors, (sample, feature, factor) (reference gibbs.py:63-73)."""
        return self._stack_samples(lambda fm: fm.V)

    def _predict_core(self, X: Optional[ArrayLike], X_rel: Sequence[RelationBlock] = (),
                      n_workers: Optional[int] = None) -> DenseArray:
        """reference gibbs.py:75-91"""
        predictor = self._fetch_predictor()
        X_rel = list(X_rel)
        X = _main_table(X, check_data_consistency(X, X_rel))
        if n_workers is None:
            return predictor.predict(X, X_rel)
        return predictor.predict_parallel(X, X_rel, n_workers)

    @classmethod
    def _train_core(cls, rank, init_stdev, X, X_rel, y, random_seed, config, callback):
        return create_train_fm(rank, init_stdev, X, X_rel, y, random_seed, config, callback)

    def get_hyper_trace(self):
        """One row per sweep: alpha, mu_w[g], lambda_w[g], mu_V[g,r], lambda_V[g,r]
        (reference gibbs.py:109-142)."""
        import pandas as pd

        if (self.n_groups_ is None) or (self.history_ is None):
            raise RuntimeError("Sampler not run yet.")
        G, K = self.n_groups_, self.rank
        columns = (["alpha"] + [f"mu_w[{g}]" for g in range(G)] + [f"lambda_w[{g}]" for g in range(G)]
                   + [f"mu_V[{g},{r}]" for g in range(G) for r in range(K)]
                   + [f"lambda_V[{g},{r}]" for g in range(G) for r in range(K)])
        rows = [
            np.concatenate([[h.alpha], h.mu_w.ravel(), h.lambda_w.ravel(), h.mu_V.ravel(),
                            h.lambda_V.ravel()])
            for h in self.history_.hypers
        ]
        return pd.DataFrame(np.vstack(rows), columns=columns)

    def _fit_gibbs(self, X, y, X_rel, X_test, y_test, X_rel_test, n_iter, n_kept_samples, grouping,
                   group_shapes, callback, config_builder, **extra):
        self._fit(X, y, X_rel=X_rel, X_test=X_test, X_rel_test=X_rel_test, y_test=y_test,
                  n_iter=n_iter, n_kept_samples=n_kept_samples, grouping=grouping,
                  callback=callback, group_shapes=group_shapes, config_builder=config_builder,
                  **extra)
        return self


class MyFMGibbsRegressor(RegressorMixin[FM, FMHyperParameters], MyFMGibbsBase):
    """Bayesian Factorization Machine for regression."""

    def fit(self, X: ArrayLike, y: np.ndarray, X_rel: Sequence[RelationBlock] = (),
            X_test: Optional[ArrayLike] = None, y_test: Optional[np.ndarray] = None,
            X_rel_test: Sequence[RelationBlock] = (), n_iter: int = 100,
            n_kept_samples: Optional[int] = None, grouping: Optional[List[int]] = None,
            group_shapes: Optional[List[int]] = None, callback=None,
            config_builder: Optional[ConfigBuilder] = None) -> "MyFMGibbsRegressor":
        return self._fit_gibbs(X, y, X_rel, X_test, y_test, X_rel_test, n_iter, n_kept_samples,
                               grouping, group_shapes, callback, config_builder)

    fit.__doc__ = _FIT_DOC

    def predict(self, X: Optional[ArrayLike], X_rel: Sequence[RelationBlock] = (),
                n_workers: Optional[int] = None) -> DenseArray:
        """Posterior predictive mean, one value per row (reference gibbs.py:219-240)."""
        return self._predict_core(X, X_rel, n_workers=n_workers)


class MyFMGibbsClassifier(ClassifierMixin[FM, FMHyperParameters], MyFMGibbsBase):
    """Bayesian Factorization Machine for binary classification (probit link)."""

    def fit(self, X: ArrayLike, y: np.ndarray, X_rel: Sequence[RelationBlock] = (),
            X_test: Optional[ArrayLike] = None, y_test: Optional[np.ndarray] = None,
            X_rel_test: Sequence[RelationBlock] = (), n_iter: int = 100,
            n_kept_samples: Optional[int] = None, grouping: Optional[List[int]] = None,
            group_shapes: Optional[List[int]] = None, callback=None,
            config_builder: Optional[ConfigBuilder] = None) -> "MyFMGibbsClassifier":
        return self._fit_gibbs(X, y, X_rel, X_test, y_test, X_rel_test, n_iter, n_kept_samples,
                               grouping, group_shapes, callback, config_builder)

    fit.__doc__ = _FIT_DOC

    def predict(self, X: Optional[ArrayLike], X_rel: Sequence[RelationBlock] = (),
                n_workers: Optional[int] = None) -> np.ndarray:
        """Class decision at probability 0.5 (reference gibbs.py:323-345)."""
        return self.predict_proba(X, X_rel, n_workers=n_workers) > 0.5

    def predict_proba(self, X: Optional[ArrayLike], X_rel: Sequence[RelationBlock] = (),
                      n_workers: Optional[int] = None) -> DenseArray:
        """Posterior mean of P(y = 1) (reference gibbs.py:347-371)."""
        return self._predict_core(X, X_rel, n_workers=n_workers)


class MyFMOrderedProbit(MyFMGibbsBase):
    """Bayesian Factorization Machine for ordinal regression (ordered probit)."""

    @property
    def _task_type(self) -> TaskType:
        return TaskType.ORDERED

    def fit(self, X: ArrayLike, y: np.ndarray, X_rel: Sequence[RelationBlock] = (),
            X_test: Optional[ArrayLike] = None, y_test: Optional[np.ndarray] = None,
            X_rel_test: Sequence[RelationBlock] = (), n_iter: int = 100,
            n_kept_samples: Optional[int] = None, grouping: Optional[List[int]] = None,
            group_shapes: Optional[List[int]] = None, callback=None,
            callback_default_freq: int = 5) -> "MyFMOrderedProbit":
        # one cut-point group holding every row (reference gibbs.py:427-432)
        y = np.asarray(y)
        builder = ConfigBuilder()
        groups = [(int(y.max() + 1), np.arange(y.shape[0]))]
        self.n_cutpoint_groups = len(groups)
        builder.set_cutpoint_groups(groups)
        return self._fit_gibbs(X, y, X_rel, X_test, y_test, X_rel_test, n_iter, n_kept_samples,
                               grouping, group_shapes, callback, builder,
                               callback_default_freq=callback_default_freq)

    fit.__doc__ = _FIT_DOC

    def _prepare_prediction_for_test(self, fm: FM, X: ArrayLike, X_rel) -> np.ndarray:
        return fm.oprobit_predict_proba(sps.csr_matrix(X, dtype=np.float64), X_rel, 0)

    def _process_y(self, y: np.ndarray) -> np.ndarray:
        assert y.min() >= 0
        return y.astype(np.float64)

    def _measure_score(self, prediction: np.ndarray, y: np.ndarray) -> Dict[str, float]:
        picked = prediction[np.arange(prediction.shape[0]), y.astype(np.int64)]
        return OrderedDict(
            accuracy=float((np.argmax(prediction, axis=1) == y).mean()),
            log_loss=float(-np.log(picked + 1e-15).mean()),
        )

    def _status_report(self, fm: FM, hyper: FMHyperParameters) -> str:
        log_str = "w0 = {:.2f}, ".format(fm.w0)
        if len(fm.cutpoints) == 1:
            log_str += "cutpoint = {} ".format(["{:.3f}".format(c) for c in list(fm.cutpoints[0])])
        return log_str

    def predict_proba(self, X: ArrayLike, X_rel: Sequence[RelationBlock] = (),
                      n_workers: Optional[int] = None) -> np.ndarray:
        """Posterior mean of the class probabilities, (row, class) (reference gibbs.py:478-509)."""
        predictor = self._fetch_predictor()
        X_rel = list(X_rel)
        X = _main_table(X, check_data_consistency(X, X_rel))
        if X.dtype != REAL:
            X = X.astype(REAL)
        return predictor.predict_parallel_oprobit(X, X_rel, n_workers or 1, 0)

    def predict(self, X: ArrayLike, X_rel: Sequence[RelationBlock] = ()) -> np.ndarray:
        """Most probable class (reference gibbs.py:511-532)."""
        return self.predict_proba(X, X_rel=X_rel).argmax(axis=1)

    @property
    def cutpoint_samples(self) -> Optional[DenseArray]:
        """Kept samples of the cut-points (reference gibbs.py:534-543)."""
        return self._stack_samples(lambda fm: fm.cutpoints[0])
