"""Estimator plumbing shared by the Gibbs models: argument normalisation, config assembly and the
progress-bar-wrapped callback.  Host-side restatement (numpy >= 2 clean) of the behaviour of the
reference's ``src/myfm/base.py``; the training itself happens in ``_myfm.create_train_fm`` on the GPU.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Callable, Dict, Generic, List, Optional, Sequence, Tuple, TypeVar, Union

import numpy as np
from scipy import sparse as sps
from scipy import special
from tqdm import tqdm

from . import _myfm
from ._myfm import ConfigBuilder, FMLearningConfig, RelationBlock, TaskType

REAL = np.float64
ArrayLike = Union[np.ndarray, sps.csr_matrix]
DenseArray = np.ndarray
BinaryClassificationTarget = np.ndarray
ClassIndexArray = np.ndarray

FM = TypeVar("FM")
Hyper = TypeVar("Hyper")
Predictor = TypeVar("Predictor")
History = TypeVar("History")
CallBackType = Callable[[int, FM, Hyper], bool]

_PRIOR_KEYS = ("alpha_0", "beta_0", "gamma_0", "mu_0", "reg_0", "fit_w0", "fit_linear")


def std_cdf(x: DenseArray) -> DenseArray:
    """Standard normal CDF (reference base.py:41-43)."""
    return (1 + special.erf(x * np.sqrt(0.5))) / 2


def check_data_consistency(X: Optional[ArrayLike], X_rel: Sequence[RelationBlock]) -> int:
    """Number of cases described by (X, X_rel); reference base.py:46-61."""
    if X_rel:
        sizes = {rel.mapper_size for rel in X_rel}
        if len(sizes) > 1:
            raise ValueError("Inconsistent case size for X_rel.")
        (n,) = sizes
        if X is not None and X.shape[0] != n:
            raise ValueError("X and X_rel have different shape.")
        return n
    if X is None:
        raise ValueError("At least X or X_rel must be provided.")
    return int(X.shape[0])


def _main_table(X: Optional[ArrayLike], n_rows: int) -> sps.csr_matrix:
    """`None` stands for a main table without columns (reference base.py:230-233)."""
    return sps.csr_matrix((n_rows, 0), dtype=REAL) if X is None else sps.csr_matrix(X)


class MyFMBase(Generic[FM, Hyper, Predictor, History], ABC):
    """Bayesian Factorization Machine estimator base (reference base.py:71-351)."""

    @classmethod
    @abstractmethod
    def _train_core(cls, rank: int, init_stdev: float, X: sps.csr_matrix, X_rel: List[RelationBlock],
                    y: np.ndarray, random_seed: int, config: FMLearningConfig,
                    callback: Callable[[int, FM, Hyper, History], bool]) -> Tuple[Predictor, History]:
        raise NotImplementedError("not implemented")

    @property
    @abstractmethod
    def _task_type(self) -> TaskType:
        raise NotImplementedError("must be specified in child")

    def __init__(self, rank: int, init_stdev: float = 0.1, random_seed: int = 42,
                 alpha_0: float = 1.0, beta_0: float = 1.0, gamma_0: float = 1.0, mu_0: float = 0.0,
                 reg_0: float = 1.0, fit_w0: bool = True, fit_linear: bool = True):
        """
        rank: number of factors.  init_stdev: weights start as Normal(0, init_stdev**2).
        random_seed: seed of the whole chain.  alpha_0, beta_0: Gamma(alpha_0/2, beta_0/2) prior of
        alpha, lambda_w, lambda_V.  gamma_0, mu_0: Normal(mu_0, 1/gamma_0) prior of mu_w, mu_V.
        reg_0: inverse prior variance of w0.  fit_w0 / fit_linear: whether to fit the bias /
        the linear coefficients.
        """
        self.rank = rank
        self.init_stdev = init_stdev
        self.random_seed = random_seed
        self.alpha_0, self.beta_0, self.gamma_0, self.mu_0 = alpha_0, beta_0, gamma_0, mu_0
        self.reg_0 = reg_0
        self.fit_w0, self.fit_linear = fit_w0, fit_linear
        self.predictor_: Optional[Predictor] = None
        self.history_: Optional[History] = None
        self.n_groups_: Optional[int] = None

    def __str__(self) -> str:
        return ("{}(init_stdev={}, alpha_0={}, beta_0={}, gamma_0={}, mu_0={}, reg_0={})").format(
            self.__class__.__name__, self.init_stdev, self.alpha_0, self.beta_0, self.gamma_0,
            self.mu_0, self.reg_0)

    # -- callbacks ---------------------------------------------------------------------------
    def _create_default_callback(self, callback_default_freq: int, do_test: bool,
                                 X_test: Optional[sps.csr_matrix] = None,
                                 X_rel_test: Sequence[RelationBlock] = (),
                                 y_test: Optional[np.ndarray] = None):
        """Status line every `callback_default_freq` sweeps (reference base.py:187-205)."""

        def callback(i: int, fm: FM, hyper: Hyper, history: History) -> Tuple[bool, Optional[str]]:
            if i % callback_default_freq:
                return False, None
            log_str = self._status_report(fm, hyper)
            if do_test:
                pred_this = self._prepare_prediction_for_test(fm, X_test, X_rel_test)
                for key, metric in self._measure_score(pred_this, y_test).items():
                    log_str += " {}_this: {:.2f}".format(key, metric)
            return False, log_str

        return callback

    # -- fit ---------------------------------------------------------------------------------
    def _fit(self, X: Optional[ArrayLike], y: np.ndarray, X_rel: Sequence[RelationBlock] = (),
             X_test: Optional[ArrayLike] = None, y_test: Optional[np.ndarray] = None,
             X_rel_test: Sequence[RelationBlock] = (), n_iter: int = 100,
             n_kept_samples: Optional[int] = None, grouping: Optional[List[int]] = None,
             group_shapes: Optional[List[int]] = None, callback=None,
             config_builder: Optional[ConfigBuilder] = None, callback_default_freq: int = 10) -> None:
        """Assemble the config and run the chain (reference base.py:207-323)."""
        builder = ConfigBuilder() if config_builder is None else config_builder
        X_rel, X_rel_test = list(X_rel), list(X_rel_test)

        X = _main_table(X, check_data_consistency(X, X_rel))
        y = np.asarray(y)
        assert X.shape[0] == y.shape[0]
        dim_all = X.shape[1] + sum(rel.feature_size for rel in X_rel)

        if n_kept_samples is None:
            n_kept_samples = min(max(n_iter - 5, 5), n_iter)
        else:
            assert n_iter >= n_kept_samples

        for key in _PRIOR_KEYS:
            getattr(builder, "set_" + key)(getattr(self, key))

        if grouping is None and group_shapes is not None:
            grouping = np.repeat(np.arange(len(group_shapes)), group_shapes).tolist()
        if grouping is None:
            self.n_groups_ = 1
            builder.set_identical_groups(dim_all)
        else:
            assert dim_all == len(grouping)
            self.n_groups_ = len(set(grouping))
            builder.set_group_index(grouping)

        do_test = X_test is not None or bool(X_rel_test)
        if do_test:
            if y_test is None:
                raise RuntimeError("Must specify both (X_test or X_rel_test) and y_test.")
            n_test = check_data_consistency(X_test, X_rel_test)
            assert n_test == y_test.shape[0]
            X_test = _main_table(X_test, n_test)
        elif y_test is not None:
            raise RuntimeError("Must specify both (X_test or X_rel_test) and y_test.")

        builder.set_n_iter(n_iter).set_n_kept_samples(n_kept_samples)
        if X.dtype != REAL:
            X = X.astype(REAL)
        y = self._process_y(y)
        builder.set_task_type(self._task_type)
        config = builder.build()

        user_callback = callback if callback is not None else self._create_default_callback(
            callback_default_freq=callback_default_freq, do_test=do_test, X_test=X_test,
            X_rel_test=X_rel_test, y_test=y_test)

        with tqdm(total=n_iter) as pbar:

            def wrapped(i: int, fm: FM, hyper: Hyper, history: History) -> bool:
                should_stop, message = user_callback(i, fm, hyper, history)
                if message is not None:
                    pbar.set_description(message)
                pbar.update(1)
                return should_stop

            # the default callback touches neither `fm` nor the device between its status lines: the next
            # sweep may start before it runs (create_train_fm); a user's callback says so itself (`observer`)
            if callback is None:
                wrapped.observer = lambda i: bool(i % callback_default_freq)
            else:
                wrapped.observer = getattr(callback, "observer", False)

            self.predictor_, self.history_ = self._train_core(
                self.rank, self.init_stdev, X, X_rel, y, self.random_seed, config, wrapped)

    # -- hooks -------------------------------------------------------------------------------
    @abstractmethod
    def _status_report(self, fm: FM, hyper: Hyper) -> str:
        raise NotImplementedError("must be implemented")

    @abstractmethod
    def _prepare_prediction_for_test(self, fm: FM, X: Optional[ArrayLike],
                                     X_rel: Sequence[RelationBlock]) -> np.ndarray:
        raise NotImplementedError("must be implemented")

    def _process_y(self, y: np.ndarray) -> DenseArray:
        return y.astype(np.float64)

    @abstractmethod
    def _measure_score(self, prediction: DenseArray, y: np.ndarray) -> Dict[str, float]:
        raise NotImplementedError("")

    def _fetch_predictor(self) -> Predictor:
        if self.predictor_ is None:
            raise RuntimeError("Predictor called before fit.")
        return self.predictor_


class RegressorMixin(Generic[FM, Hyper]):
    """reference base.py:353-374"""

    @property
    def _task_type(self) -> TaskType:
        return TaskType.REGRESSION

    def _prepare_prediction_for_test(self, fm, X, X_rel) -> np.ndarray:
        return fm.predict_score(X, X_rel)

    def _status_report(self, fm, hyper) -> str:
        return "alpha = {:.2f} w0 = {:.2f} ".format(hyper.alpha, fm.w0)

    def _measure_score(self, prediction: np.ndarray, y: np.ndarray) -> Dict[str, float]:
        err = y - prediction
        return OrderedDict(rmse=float(np.sqrt(np.mean(err ** 2))), mae=float(np.mean(np.abs(err))))


class ClassifierMixin(Generic[FM, Hyper], ABC):
    """reference base.py:377-399; labels {0,1} -> targets {-1,+1}."""

    @property
    def _task_type(self) -> TaskType:
        return TaskType.CLASSIFICATION

    def _prepare_prediction_for_test(self, fm, X, X_rel) -> np.ndarray:
        return std_cdf(fm.predict_score(X, X_rel))

    def _process_y(self, y: np.ndarray) -> np.ndarray:
        return y.astype(np.float64) * 2 - 1

    def _measure_score(self, prediction: np.ndarray, y: np.ndarray) -> Dict[str, float]:
        gt = y > 0
        lp, l1mp = np.log(prediction + 1e-15), np.log(1 - prediction + 1e-15)
        return OrderedDict(
            ll=float((-lp.dot(gt) - l1mp.dot(~gt)) / max(1, prediction.shape[0])),
            accuracy=float(np.mean((prediction >= 0.5) == gt)),
        )

    def _status_report(self, fm, hyper) -> str:
        return "w0 = {:.2f} ".format(fm.w0)
