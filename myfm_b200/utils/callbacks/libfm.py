"""LibFM-style per-iteration evaluation callbacks (reference src/myfm/utils/callbacks/libfm.py).

Every sweep the callback scores the held-out set with the CURRENT sample, keeps running means
(over all sweeps, and over all but the first five) and appends the metrics to ``result_trace``.
When the `fm` handed in is the engine's live, device-resident sample (what `fit()` passes), the
whole step runs on the device (`_myfm._DeviceEvaluator` -> csrc/eval_device.cuh): forward pass, link,
both running sums and the metric reductions against a test matrix that stays in HBM; nine scalars
come back per sweep, and `predictions` / `prediction_all_but_5` are fetched only when read.  A plain
host `FM` takes the numpy path of the reference.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
from scipy import sparse as sps

from ..._myfm import FM, FMHyperParameters, LearningHistory, RelationBlock, TaskType, _DeviceEvaluator, _LiveFM
from ...base import REAL, ArrayLike, check_data_consistency, std_cdf

_BURN_IN = 5


class LibFMLikeCallbackBase(ABC):
    def __init__(self, n_iter: int, X_test: Optional[ArrayLike], X_rel_test: List[RelationBlock],
                 y_test: np.ndarray, trace_path: Optional[str] = None):
        self.n_test_data = check_data_consistency(X_test, X_rel_test)
        self.n_iter = n_iter
        if X_test is None:
            self.X_test = sps.csr_matrix((self.n_test_data, 0), dtype=REAL)
        else:
            self.X_test = sps.csr_matrix(X_test, dtype=REAL)
        self.X_rel_test = X_rel_test
        self.y_test: np.ndarray = y_test
        self.result_trace: List[Dict[str, float]] = []
        self.trace_path = trace_path
        self.n_samples = 0

    def _accumulate(self, i: int, this: np.ndarray):
        """Adds this sweep's prediction to the running sums; returns (mean, mean_all_but_5|None)."""
        self.predictions += this
        self.n_samples += 1
        mean = self.predictions / self.n_samples
        late = None
        if i >= _BURN_IN:
            self.prediction_all_but_5 += this
            late = self.prediction_all_but_5 / (i + 1 - _BURN_IN)
        return mean, late

    # ---- device path -----------------------------------------------------------------------------
    _device: Optional[_DeviceEvaluator] = None
    _host_predictions: np.ndarray
    _host_all_but_5: np.ndarray

    def _device_kwargs(self) -> Dict:
        raise NotImplementedError

    def _device_terms(self, i: int, fm: FM, cutpoints=None) -> Optional[np.ndarray]:
        """The nine metric sums of this step from the device, or None when `fm` is a host object (or a host
        step has already been taken: the running sums then live in numpy)."""
        if not isinstance(fm, _LiveFM) or (self._device is None and self.n_samples > 0):
            return None
        if self._device is None:
            self._device = _DeviceEvaluator(fm, self.X_test, self.X_rel_test, self.y_test, **self._device_kwargs())
        self.n_samples += 1
        return self._device.step(fm, i, cutpoints)

    # the running sums live on the device while the device path is in use; reading them copies them
    @property
    def predictions(self) -> np.ndarray:
        return self._device.sums()[0] if self._device is not None else self._host_predictions

    @predictions.setter
    def predictions(self, value: np.ndarray) -> None:
        self._host_predictions = value

    @property
    def prediction_all_but_5(self) -> np.ndarray:
        return self._device.sums()[1] if self._device is not None else self._host_all_but_5

    @prediction_all_but_5.setter
    def prediction_all_but_5(self, value: np.ndarray) -> None:
        self._host_all_but_5 = value

    @abstractmethod
    def _measure_score(self, i: int, fm: FM, hyper: FMHyperParameters) -> Tuple[str, Dict[str, float]]:
        raise NotImplementedError("must be implemented")

    def __call__(self, i: int, fm: FM, hyper: FMHyperParameters,
                 history: LearningHistory) -> Tuple[bool, Optional[str]]:
        description, trace_result = self._measure_score(i, fm, hyper)
        self.result_trace.append(trace_result)
        if self.trace_path is not None:
            import pandas as pd

            pd.DataFrame(self.result_trace).to_csv(self.trace_path, index=False)
        return False, description


class RegressionCallback(LibFMLikeCallbackBase):
    def __init__(self, n_iter: int, X_test: Optional[ArrayLike], y_test: np.ndarray,
                 X_rel_test: List[RelationBlock] = [], clip_min: Optional[float] = None,
                 clip_max: Optional[float] = None, trace_path: Optional[str] = None):
        super().__init__(n_iter, X_test, X_rel_test, y_test, trace_path=trace_path)
        self.predictions = np.zeros((self.n_test_data,), dtype=np.float64)
        self.prediction_all_but_5 = np.zeros((self.n_test_data,), dtype=np.float64)
        self.clip_min, self.clip_max = clip_min, clip_max

    def clip_value(self, arr: np.ndarray) -> None:
        if self.clip_min is not None:
            arr[arr <= self.clip_min] = self.clip_min
        if self.clip_max is not None:
            arr[arr >= self.clip_max] = self.clip_max

    def _rmse(self, pred: np.ndarray) -> float:
        return float(((self.y_test - pred) ** 2).mean() ** 0.5)

    def _device_kwargs(self):
        return dict(task=TaskType.REGRESSION, clip_min=self.clip_min, clip_max=self.clip_max)

    def _measure_score(self, i, fm, hyper):
        terms = self._device_terms(i, fm)
        if terms is not None:
            n = max(1, self.n_test_data)
            rmse, rmse_this = (terms[0] / n) ** 0.5, (terms[1] / n) ** 0.5
            rmse_late = (terms[2] / n) ** 0.5 if i >= _BURN_IN else float("nan")
        else:
            score = fm.predict_score(self.X_test, self.X_rel_test)
            mean, late = self._accumulate(i, score)
            self.clip_value(mean)
            rmse_late = float("nan")
            if late is not None:
                self.clip_value(late)
                rmse_late = self._rmse(late)
            rmse, rmse_this = self._rmse(mean), self._rmse(score)
        description = "alpha={0:.4f}, rmse_mean={1:.4f}, rmse_this={2:.4f}, rmse_all_but_5={3:.4f}".format(
            hyper.alpha, rmse, rmse_this, rmse_late)
        return description, OrderedDict(
            [("alpha", hyper.alpha), ("rmse", rmse), ("rmse_this", rmse_this), ("rmse_all_but_5", rmse_late)])


class ClassificationCallback(LibFMLikeCallbackBase):
    def __init__(self, n_iter: int, X_test: Optional[ArrayLike], y_test: np.ndarray,
                 X_rel_test: List[RelationBlock] = [], eps: Optional[float] = 1e-15,
                 trace_path: Optional[str] = None):
        super().__init__(n_iter, X_test, X_rel_test, y_test, trace_path=trace_path)
        self.predictions = np.zeros((self.n_test_data,), dtype=np.float64)
        self.prediction_all_but_5 = np.zeros((self.n_test_data,), dtype=np.float64)
        self.eps = eps

    def clip_value(self, arr: np.ndarray) -> None:
        if self.eps is not None:
            arr[arr <= self.eps] = self.eps
            arr[arr >= (1 - self.eps)] = 1 - self.eps

    def _log_loss(self, p: np.ndarray) -> float:
        return -float(np.log(p[self.y_test == 1]).sum() + np.log(1 - p[self.y_test == 0]).sum())

    def _accuracy(self, p: np.ndarray) -> float:
        return float((self.y_test == (p >= 0.5)).mean())

    def _device_kwargs(self):
        return dict(task=TaskType.CLASSIFICATION, eps=self.eps)

    def _measure_score(self, i, fm, hyper):
        terms = self._device_terms(i, fm)
        if terms is not None:
            n = max(1, self.n_test_data)
            ll, ll_this, acc, acc_this = terms[0], terms[1], terms[3] / n, terms[4] / n
            ll_late, acc_late = (terms[2], terms[5] / n) if i >= _BURN_IN else (float("nan"), float("nan"))
        else:
            prob_this = std_cdf(fm.predict_score(self.X_test, self.X_rel_test))
            mean, late = self._accumulate(i, prob_this)
            self.clip_value(mean)
            ll_late = acc_late = float("nan")
            if late is not None:
                self.clip_value(late)
                ll_late, acc_late = self._log_loss(late), self._accuracy(late)
            ll, acc = self._log_loss(mean), self._accuracy(mean)
            ll_this, acc_this = self._log_loss(prob_this), self._accuracy(prob_this)
        description = "ll_mean={0:.4f}, ll_this={1:.4f}, ll_all_but_5={2:.4f}".format(ll, ll_this, ll_late)
        return description, OrderedDict([
            ("log_loss", ll), ("log_loss_this", ll_this), ("log_loss_all_but_5", ll_late),
            ("accuracy", acc), ("accuracy_this", acc_this), ("accuracy_all_but_5", acc_late)])


class OrderedProbitCallback(LibFMLikeCallbackBase):
    def __init__(self, n_iter: int, X_test: Optional[ArrayLike], y_test: np.ndarray, n_class: int,
                 X_rel_test: List[RelationBlock] = [], eps: Optional[float] = 1e-15,
                 trace_path: Optional[str] = None):
        super().__init__(n_iter, X_test, X_rel_test, y_test, trace_path=trace_path)
        self.predictions = np.zeros((self.n_test_data, n_class), dtype=np.float64)
        self.prediction_all_but_5 = np.zeros((self.n_test_data, n_class), dtype=np.float64)
        self.n_class = n_class
        self.eps = eps
        self.y_test = self.y_test.astype(np.int32)
        assert (self.y_test.min() >= 0) and (self.y_test.max() <= (self.n_class - 1))

    def _log_loss(self, p: np.ndarray) -> float:
        ps = p[np.arange(self.y_test.shape[0]), self.y_test].copy()
        ps[ps <= self.eps] = self.eps
        return -float(np.log(ps).sum())

    def _accuracy(self, p: np.ndarray) -> float:
        return float((self.y_test == p.argmax(axis=1)).mean())

    def _rmse(self, p: np.ndarray) -> float:
        return float(((self.y_test - p.dot(np.arange(self.n_class))) ** 2).mean()) ** 0.5

    def _device_kwargs(self):
        return dict(task=TaskType.ORDERED, n_class=self.n_class, eps=self.eps)

    def _measure_score(self, i, fm, hyper):
        terms = self._device_terms(i, fm, fm.cutpoints[0]) if isinstance(fm, _LiveFM) else None
        if terms is not None:
            n = max(1, self.n_test_data)
            ll, ll_this, acc, acc_this = terms[0], terms[1], terms[3] / n, terms[4] / n
            rmse, rmse_this = (terms[6] / n) ** 0.5, (terms[7] / n) ** 0.5
            ll_late = acc_late = rmse_late = float("nan")
            if i >= _BURN_IN:
                ll_late, acc_late, rmse_late = terms[2], terms[5] / n, (terms[8] / n) ** 0.5
        else:
            prob_this = fm.oprobit_predict_proba(self.X_test, self.X_rel_test, 0)
            mean, late = self._accumulate(i, prob_this)
            ll_late = acc_late = rmse_late = float("nan")
            if late is not None:
                ll_late, acc_late, rmse_late = self._log_loss(late), self._accuracy(late), self._rmse(late)
            ll, acc, rmse = self._log_loss(mean), self._accuracy(mean), self._rmse(mean)
            ll_this, acc_this, rmse_this = self._log_loss(prob_this), self._accuracy(prob_this), self._rmse(prob_this)
        description = "ll_mean={0:.4f}, ll_this={1:.4f}, ll_all_but_5={2:.4f}".format(ll, ll_this, ll_late)
        return description, OrderedDict([
            ("log_loss", ll), ("log_loss_this", ll_this), ("log_loss_all_but_5", ll_late),
            ("accuracy", acc), ("accuracy_this", acc_this), ("accuracy_all_but_5", acc_late),
            ("rmse", rmse), ("rmse_this", rmse_this), ("rmse_all_but_5", rmse_late)])
