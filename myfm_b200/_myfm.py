"""Host-side mirror of the reference's pybind11 extension ``myfm._myfm``
(reference cpp_source/declare_module.hpp:67-404; stub src/myfm/_myfm.pyi), implemented over the
C ABI of the CUDA engine (include/myfm_b200.h).

Same names, argument meaning, return types, pickling layouts and error behaviour as the
extension it replaces, so ``base.py`` / ``gibbs.py`` / the callbacks read exactly like the
reference's.  State containers (FM, FMHyperParameters, ...) are plain numpy holders; everything
that touches the training data or computes a prediction goes to the GPU through ``_lib``.
"""
from __future__ import annotations

import ctypes as C
import os
import enum
import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse as sps
from scipy import special

from . import _lib
from .options import engine_options, get_options

__all__ = [
    "TaskType", "FMLearningConfig", "ConfigBuilder", "RelationBlock", "FM", "FMHyperParameters",
    "Predictor", "LearningHistory", "FMTrainer", "create_train_fm", "create_train_vfm",
    "mean_var_truncated_normal_left", "mean_var_truncated_normal_right",
]


class TaskType(enum.IntEnum):
    """declare_module.hpp:88-91"""

    REGRESSION = 0
    CLASSIFICATION = 1
    ORDERED = 2


# ------------------------------------------------------------------------------------------------
# RelationBlock — declare_module.hpp:95-137, include/myfm/definitions.hpp:30-52
# ------------------------------------------------------------------------------------------------
class RelationBlock:
    """The RelationBlock Class.

    Parameters
    ----------
    original_to_block: List[int]
        describes which entry points to to which row of the data (second argument).
    data: scipy.sparse.csr_matrix[float64]
        describes repeated pattern.

    Note
    -----
    The entries of `original_to_block` must be in the [0, data.shape[0]-1].
    """

    def __init__(self, original_to_block, data) -> None:
        omap = np.asarray(original_to_block)
        if omap.ndim != 1:
            raise TypeError("original_to_block must be a one-dimensional sequence of int")
        if omap.size and not np.issubdtype(omap.dtype, np.integer):
            if not np.all(np.equal(np.mod(omap, 1), 0)):
                raise TypeError("original_to_block must contain integers")
        if omap.size and omap.min() < 0:
            raise TypeError("original_to_block must contain non-negative integers")
        self._map = np.ascontiguousarray(omap, dtype=np.int64)
        X = sps.csr_matrix(data)
        if X.dtype != np.float64:
            X = X.astype(np.float64)
        self._data = X
        if self._map.size and self._map.max() >= X.shape[0]:
            raise RuntimeError("index mapping points to non-existing row.")

    @property
    def original_to_block(self) -> List[int]:
        return self._map.tolist()

    @property
    def data(self) -> sps.csr_matrix:
        return self._data

    @property
    def mapper_size(self) -> int:
        return int(self._map.shape[0])

    @property
    def block_size(self) -> int:
        return int(self._data.shape[0])

    @property
    def feature_size(self) -> int:
        return int(self._data.shape[1])

    def __repr__(self) -> str:
        return "<RelationBlock with mapper size = {}, block data size = {}, feature size = {}>".format(
            self.mapper_size, self.block_size, self.feature_size)

    def __getstate__(self):
        return (self._map.tolist(), self._data)

    def __setstate__(self, state) -> None:
        if len(state) != 2:
            raise RuntimeError("invalid state for Relationblock.")
        self.__init__(state[0], state[1])


# ------------------------------------------------------------------------------------------------
# FMLearningConfig / ConfigBuilder — include/myfm/FMLearningConfig.hpp, declare_module.hpp:139-156
# ------------------------------------------------------------------------------------------------
class FMLearningConfig:
    """Opaque, validated learning configuration (built by ConfigBuilder.build)."""

    def __init__(self, **fields) -> None:
        self.__dict__.update(fields)
        self._keep = []
        self.n_groups = self._validate()

    def _as_struct(self) -> _lib.Config:
        c = _lib.Config()
        for k in ("alpha_0", "beta_0", "gamma_0", "mu_0", "reg_0", "nu_oprobit", "cutpoint_scale"):
            setattr(c, k, float(getattr(self, k)))
        c.task_type = int(self.task_type)
        c.fit_w0, c.fit_linear = int(self.fit_w0), int(self.fit_linear)
        c.n_iter, c.n_kept_samples = int(self.n_iter), int(self.n_kept_samples)
        gi = np.ascontiguousarray(self.group_index, dtype=np.int64)
        cg = self.cutpoint_groups
        ncls = np.asarray([g[0] for g in cg] or [0], dtype=np.int32)
        rows = [np.ascontiguousarray(g[1], dtype=np.int64) for g in cg]
        lens = np.asarray([r.shape[0] for r in rows] or [0], dtype=np.int64)
        ptrs = (C.POINTER(C.c_int64) * max(1, len(cg)))(*[_lib.ptr(r, C.c_int64) for r in rows])
        c.group_index, c.n_group_index = _lib.ptr(gi, C.c_int64), gi.shape[0]
        c.n_cutpoint_groups = len(cg)
        c.cutpoint_n_class = _lib.ptr(ncls, C.c_int32)
        c.cutpoint_index = C.cast(ptrs, C.POINTER(C.POINTER(C.c_int64)))
        c.cutpoint_index_len = _lib.ptr(lens, C.c_int64)
        self._keep = [gi, ncls, rows, lens, ptrs]
        return c

    def _validate(self) -> int:
        n_groups = C.c_int32(0)
        _lib.check(_lib.lib().myfm_config_validate(C.byref(self._as_struct()), C.byref(n_groups)))
        return int(n_groups.value)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_keep"] = []
        return d


class ConfigBuilder:
    """FMLearningConfig::Builder (FMLearningConfig.hpp:92-201); every setter returns self."""

    def __init__(self) -> None:
        self.alpha_0 = 1.0
        self.beta_0 = 1.0
        self.gamma_0 = 1.0
        self.mu_0 = 1.0
        self.reg_0 = 1.0
        self.n_iter = 100
        self.n_kept_samples = 10
        self.task_type = TaskType.REGRESSION
        self.nu_oprobit = 5.0
        self.fit_w0 = True
        self.fit_linear = True
        self.group_index: List[int] = []
        self.cutpoint_scale = 10.0
        self.cutpoint_groups: List[Tuple[int, List[int]]] = []

    def set_alpha_0(self, arg: float) -> "ConfigBuilder":
        self.alpha_0 = float(arg)
        return self

    def set_beta_0(self, arg: float) -> "ConfigBuilder":
        self.beta_0 = float(arg)
        return self

    def set_gamma_0(self, arg: float) -> "ConfigBuilder":
        self.gamma_0 = float(arg)
        return self

    def set_mu_0(self, arg: float) -> "ConfigBuilder":
        self.mu_0 = float(arg)
        return self

    def set_reg_0(self, arg: float) -> "ConfigBuilder":
        self.reg_0 = float(arg)
        return self

    def set_n_iter(self, arg: int) -> "ConfigBuilder":
        self.n_iter = int(arg)
        return self

    def set_n_kept_samples(self, arg: int) -> "ConfigBuilder":
        self.n_kept_samples = int(arg)
        return self

    def set_task_type(self, arg: TaskType) -> "ConfigBuilder":
        self.task_type = TaskType(arg)
        return self

    def set_nu_oprobit(self, arg: int) -> "ConfigBuilder":
        self.nu_oprobit = float(int(arg))
        return self

    def set_fit_w0(self, arg: bool) -> "ConfigBuilder":
        self.fit_w0 = bool(arg)
        return self

    def set_fit_linear(self, arg: bool) -> "ConfigBuilder":
        self.fit_linear = bool(arg)
        return self

    def set_group_index(self, arg: Sequence[int]) -> "ConfigBuilder":
        self.group_index = np.asarray(arg, dtype=np.int64)
        return self

    def set_identical_groups(self, arg: int) -> "ConfigBuilder":
        self.group_index = np.zeros(int(arg), dtype=np.int64)
        return self

    def set_cutpoint_scale(self, arg: float) -> "ConfigBuilder":
        self.cutpoint_scale = float(arg)
        return self

    def set_cutpoint_groups(self, arg: Sequence[Tuple[int, Sequence[int]]]) -> "ConfigBuilder":
        self.cutpoint_groups = [(int(n), np.asarray(rows, dtype=np.int64)) for n, rows in arg]
        return self

    def build(self) -> FMLearningConfig:
        return FMLearningConfig(
            alpha_0=self.alpha_0, beta_0=self.beta_0, gamma_0=self.gamma_0, mu_0=self.mu_0,
            reg_0=self.reg_0, task_type=self.task_type, nu_oprobit=self.nu_oprobit,
            fit_w0=self.fit_w0, fit_linear=self.fit_linear, group_index=self.group_index,
            n_iter=self.n_iter, n_kept_samples=self.n_kept_samples,
            cutpoint_scale=self.cutpoint_scale, cutpoint_groups=list(self.cutpoint_groups),
        )


# ------------------------------------------------------------------------------------------------
# device datasets (a test / training design matrix resident in HBM)
# ------------------------------------------------------------------------------------------------
class _DeviceDataset:
    def __init__(self, X, relations: Sequence[RelationBlock], dtype: str, device: int) -> None:
        self._h = None
        csr = _lib.CsrHolder(X)
        rels = _lib.RelationsHolder(relations)
        h = C.c_void_p()
        _lib.check(_lib.lib().myfm_dataset_create(
            C.byref(h), C.byref(csr.struct), C.c_int32(rels.n), rels.array,
            C.c_int32(_lib.DTYPES[dtype]), C.c_int32(device)))
        self._h = h
        self.n_rows = csr.shape[0]
        self.dtype = dtype

    def __del__(self) -> None:
        if getattr(self, "_h", None) is not None and _lib is not None:  # None at interpreter exit
            _lib.lib().myfm_dataset_destroy(self._h)
            self._h = None


_dataset_cache: List[Tuple[tuple, tuple, _DeviceDataset]] = []
_DATASET_CACHE_SIZE = 4


def _as_csr(X) -> sps.csr_matrix:
    if not sps.isspmatrix_csr(X):
        X = sps.csr_matrix(X)
    return X


def _fingerprint(X: sps.csr_matrix) -> tuple:
    """Cheap content check of a cached upload: strided samples of values / indices plus their ends.
    In-place edits of a matrix between two predictions change it (a full checksum would cost as much
    as the upload the cache saves)."""
    n = X.nnz
    if n == 0:
        return (0,)
    step = max(1, n // 1024)
    return (float(X.data[::step].sum()), int(X.indices[::step].astype(np.int64).sum()), float(X.data[-1]),
            int(X.indptr[-1]))


def clear_dataset_cache() -> None:
    """Drops the cached device copies of prediction matrices (and the host references pinning them)."""
    _dataset_cache.clear()


def _device_dataset(X, relations: Sequence[RelationBlock]) -> _DeviceDataset:
    """Per-iteration callbacks predict on the same test matrix every sweep
    (reference src/myfm/utils/callbacks/libfm.py:82-113); keep the last few uploads alive, keyed by
    the identity of the underlying buffers and a sampled fingerprint of their contents
    (`clear_dataset_cache()` releases them)."""
    opts = get_options()
    X = _as_csr(X)
    key = (X.data.ctypes.data, X.indices.ctypes.data, X.indptr.ctypes.data, X.shape, X.nnz,
           tuple(id(r) for r in relations), opts.dtype, opts.device, _fingerprint(X))
    for k, _, ds in _dataset_cache:
        if k == key:
            return ds
    ds = _DeviceDataset(X, relations, opts.dtype, opts.device)
    _dataset_cache.append((key, (X, tuple(relations)), ds))  # pins the buffers the key names
    if len(_dataset_cache) > _DATASET_CACHE_SIZE:
        _dataset_cache.pop(0)
    return ds


def _check_relations(X: sps.csr_matrix, relations: Sequence[RelationBlock]) -> int:
    """check_row_consistency_return_column, include/myfm/util.hpp:147-165"""
    col = X.shape[1]
    for i, rel in enumerate(relations):
        if X.shape[0] != rel.mapper_size:
            raise RuntimeError("main table has size {} but the relation[{}] has size {}".format(
                X.shape[0], i, rel.mapper_size))
        col += rel.feature_size
    return col


# ------------------------------------------------------------------------------------------------
# FM — declare_module.hpp:158-192, include/myfm/FM.hpp
# ------------------------------------------------------------------------------------------------
class FM:
    def __init__(self, w0: float, w: np.ndarray, V: np.ndarray,
                 cutpoints: Optional[List[np.ndarray]] = None) -> None:
        self.w0 = float(w0)
        self.w = np.ascontiguousarray(w, dtype=np.float64)
        V = np.asarray(V, dtype=np.float64)
        self.V = np.ascontiguousarray(V.reshape(self.w.shape[0], -1))
        self.cutpoints = [np.asarray(c, dtype=np.float64) for c in (cutpoints or [])]

    def predict_score(self, X, relations: Sequence[RelationBlock] = ()) -> np.ndarray:
        """FM.hpp:47-136"""
        X = _as_csr(X)
        for rel in relations:
            if X.shape[0] != rel.mapper_size:
                raise ValueError("Relation blocks have inconsistent mapper size with case_size")
        ds = _device_dataset(X, list(relations))
        out = np.empty(X.shape[0], dtype=np.float64)
        _lib.check(_lib.lib().myfm_predict_score(
            ds._h, C.c_double(self.w0), _lib.vptr(self.w), _lib.vptr(self.V),
            C.c_int64(self.w.shape[0]), C.c_int32(self.V.shape[1]), _lib.vptr(out)))
        return out

    def oprobit_predict_proba(self, X, relations: Sequence[RelationBlock],
                              cutpoint_index: int) -> np.ndarray:
        """FM.hpp:137-162"""
        if not self.cutpoints:
            raise RuntimeError("No cutpoint available for this FM.")
        cp = self.cutpoints[cutpoint_index]  # IndexError <-> std::out_of_range
        X = _as_csr(X)
        ds = _device_dataset(X, list(relations))
        out = np.empty((X.shape[0], cp.shape[0] + 1), dtype=np.float64)
        w0s = np.asarray([self.w0], dtype=np.float64)
        cps = np.ascontiguousarray(cp, dtype=np.float64)
        _lib.check(_lib.lib().myfm_predict_oprobit_mean(
            ds._h, C.c_int32(1), _lib.vptr(w0s), _lib.vptr(self.w), _lib.vptr(self.V),
            _lib.vptr(cps), C.c_int32(cp.shape[0]), C.c_int64(self.w.shape[0]),
            C.c_int32(self.V.shape[1]), _lib.vptr(out)))
        return out

    def __repr__(self) -> str:
        return "<Factorization Machine sample with feature size = {}, rank = {}>".format(
            self.w.shape[0], self.V.shape[1])

    def __getstate__(self):
        return (self.w0, self.w, self.V, self.cutpoints)

    def __setstate__(self, state) -> None:
        if len(state) == 3:  # compatibility with earlier versions (declare_module.hpp:179-183)
            self.__init__(state[0], state[1], state[2])
        elif len(state) == 4:
            self.__init__(state[0], state[1], state[2], state[3])
        else:
            raise RuntimeError("invalid state for FM.")


class _LiveFM(FM):
    """The `fm` a per-iteration callback receives (FMTrainer.hpp:78): a view of the trainer's
    device-resident state, valid during the callback only — like the reference's raw pointer.
    Weights are fetched on first access; predict_score runs on the device-resident sample."""

    def __init__(self, trainer: "_TrainerHandle") -> None:  # noqa: super().__init__ not wanted
        self._trainer = trainer
        self._cache: Optional[Tuple[float, np.ndarray, np.ndarray, List[np.ndarray]]] = None

    def _fetch(self):
        if self._cache is None:
            self._cache = self._trainer.get_fm()
        return self._cache

    # the bias alone does not pull the whole sample over the bus
    w0 = property(lambda self: self._trainer.get_w0() if self._cache is None else self._cache[0])
    w = property(lambda self: self._fetch()[1])
    V = property(lambda self: self._fetch()[2])
    cutpoints = property(lambda self: self._fetch()[3])

    def predict_score(self, X, relations: Sequence[RelationBlock] = ()) -> np.ndarray:
        X = _as_csr(X)
        for rel in relations:
            if X.shape[0] != rel.mapper_size:
                raise ValueError("Relation blocks have inconsistent mapper size with case_size")
        ds = _device_dataset(X, list(relations))
        out = np.empty(X.shape[0], dtype=np.float64)
        _lib.check(_lib.lib().myfm_trainer_predict_score(self._trainer._h, ds._h, _lib.vptr(out)))
        return out

    def oprobit_predict_proba(self, X, relations, cutpoint_index: int) -> np.ndarray:
        return self.freeze().oprobit_predict_proba(X, relations, cutpoint_index)

    def freeze(self) -> FM:
        w0, w, V, cps = self._fetch()  # freshly downloaded arrays owned by this transient view
        self._cache = None
        return FM(w0, w, V, cps)

    def __getstate__(self):
        return self.freeze().__getstate__()

    def __reduce__(self):
        return (FM, self.freeze().__getstate__())


class _DeviceEvaluator:
    """Device half of the LibFM-style callbacks (myfm_evaluator_*): the held-out matrix, the running sums of
    the per-sweep predictions and the metric reductions stay in HBM; `step` returns nine sums."""

    def __init__(self, live: "_LiveFM", X, relations: Sequence[RelationBlock], y_test: np.ndarray, task: "TaskType",
                 n_class: int = 0, clip_min: Optional[float] = None, clip_max: Optional[float] = None,
                 eps: Optional[float] = None) -> None:
        self._h = None
        trainer = live._trainer
        with engine_options(dtype=trainer.dtype, device=trainer.device):
            self._ds = _device_dataset(_as_csr(X), list(relations))  # keeps the upload alive
        y = np.ascontiguousarray(y_test, dtype=np.float64)
        nan = float("nan")
        h = C.c_void_p()
        _lib.check(_lib.lib().myfm_evaluator_create(
            C.byref(h), self._ds._h, _lib.vptr(y), C.c_int64(y.shape[0]), C.c_int32(int(task)), C.c_int32(int(n_class)),
            C.c_double(nan if clip_min is None else clip_min), C.c_double(nan if clip_max is None else clip_max),
            C.c_double(-1.0 if eps is None else eps)))
        self._h = h
        self.n, self.width = y.shape[0], (int(n_class) if task == TaskType.ORDERED else 1)

    def step(self, live: "_LiveFM", iteration: int, cutpoints: Optional[np.ndarray] = None) -> np.ndarray:
        terms = np.zeros(9, dtype=np.float64)
        cp = None if cutpoints is None else np.ascontiguousarray(cutpoints, dtype=np.float64)
        _lib.check(_lib.lib().myfm_evaluator_step(
            self._h, live._trainer._h, C.c_int32(int(iteration)), None if cp is None else _lib.vptr(cp),
            C.c_int32(0 if cp is None else cp.shape[0]), _lib.vptr(terms)))
        return terms

    def sums(self) -> Tuple[np.ndarray, np.ndarray]:
        shape = (self.n,) if self.width == 1 else (self.n, self.width)
        total, late = np.zeros(shape, dtype=np.float64), np.zeros(shape, dtype=np.float64)
        _lib.check(_lib.lib().myfm_evaluator_get_sums(self._h, _lib.vptr(total), _lib.vptr(late)))
        return total, late

    def __del__(self) -> None:
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.lib().myfm_evaluator_destroy(self._h)
            self._h = None


class _DeviceFM(FM):
    """A kept sample (`Predictor.samples[i]`) that stays on the device: the copy the reference makes
    per kept iteration (FMTrainer.hpp:74-76) is device-to-device here.  `w` / `V` are downloaded on
    first access; prediction over such samples runs without any weight upload.  Pickles as a
    plain FM."""

    def __init__(self, handle: int, w0: float, cutpoints: List[np.ndarray], dim_all: int, rank: int,
                 dtype: str, device: int) -> None:  # noqa: super().__init__ not wanted
        self._h = C.c_void_p(handle)
        self.w0 = float(w0)
        self.cutpoints = cutpoints
        self._dim_all, self._rank, self._dtype, self._device = int(dim_all), int(rank), dtype, int(device)
        self._w: Optional[np.ndarray] = None
        self._V: Optional[np.ndarray] = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().myfm_sample_destroy(h)
            except Exception:
                pass

    def _materialize(self) -> None:
        if self._w is None:
            w = np.empty(self._dim_all, dtype=np.float64)
            V = np.empty((self._dim_all, self._rank), dtype=np.float64)
            _lib.check(_lib.lib().myfm_sample_get(self._h, None, _lib.vptr(w), _lib.vptr(V)))
            self._w, self._V = w, V

    @property
    def w(self) -> np.ndarray:
        self._materialize()
        return self._w

    @property
    def V(self) -> np.ndarray:
        self._materialize()
        return self._V

    def _on_device_for(self, opts) -> bool:
        return self._h is not None and opts.dtype == self._dtype and opts.device == self._device

    def predict_score(self, X, relations: Sequence[RelationBlock] = ()) -> np.ndarray:
        opts = get_options()
        if not self._on_device_for(opts):
            return FM(self.w0, self.w, self.V, self.cutpoints).predict_score(X, relations)
        X = _as_csr(X)
        for rel in relations:
            if X.shape[0] != rel.mapper_size:
                raise ValueError("Relation blocks have inconsistent mapper size with case_size")
        ds = _device_dataset(X, list(relations))
        out = np.empty(X.shape[0], dtype=np.float64)
        handles = (C.c_void_p * 1)(self._h)
        _lib.check(_lib.lib().myfm_predict_samples_mean(ds._h, C.c_int32(int(TaskType.REGRESSION)), C.c_int32(1),
                                                       handles, None, C.c_int32(-1), _lib.vptr(out)))
        return out

    def oprobit_predict_proba(self, X, relations, cutpoint_index: int) -> np.ndarray:
        return FM(self.w0, self.w, self.V, self.cutpoints).oprobit_predict_proba(X, relations, cutpoint_index)

    def __getstate__(self):
        return (self.w0, self.w, self.V, self.cutpoints)

    def __reduce__(self):
        return (FM, (self.w0, self.w, self.V, self.cutpoints))


# ------------------------------------------------------------------------------------------------
# FMHyperParameters — declare_module.hpp:238-261, include/myfm/HyperParams.hpp
# ------------------------------------------------------------------------------------------------
class FMHyperParameters:
    def __init__(self, alpha: float, mu_w: np.ndarray, lambda_w: np.ndarray, mu_V: np.ndarray,
                 lambda_V: np.ndarray) -> None:
        self._alpha = float(alpha)
        self._mu_w = np.asarray(mu_w, dtype=np.float64)
        self._lambda_w = np.asarray(lambda_w, dtype=np.float64)
        self._mu_V = np.asarray(mu_V, dtype=np.float64)
        self._lambda_V = np.asarray(lambda_V, dtype=np.float64)

    alpha = property(lambda self: self._alpha)
    mu_w = property(lambda self: self._mu_w)
    lambda_w = property(lambda self: self._lambda_w)
    mu_V = property(lambda self: self._mu_V)
    lambda_V = property(lambda self: self._lambda_V)

    def __getstate__(self):
        return (self._alpha, self._mu_w, self._lambda_w, self._mu_V, self._lambda_V)

    def __setstate__(self, state) -> None:
        if len(state) != 5:
            raise RuntimeError("invalid state for FMHyperParameters.")
        self.__init__(*state)


# ------------------------------------------------------------------------------------------------
# Predictor — declare_module.hpp:303-323, include/myfm/predictor.hpp
# ------------------------------------------------------------------------------------------------
class Predictor:
    def __init__(self, rank: int, feature_size: int, task_type: int) -> None:
        self._rank = int(rank)
        self._feature_size = int(feature_size)
        self._type = TaskType(int(task_type))
        self.samples: List[FM] = []

    def _check_input(self, X, relations: Sequence[RelationBlock]) -> sps.csr_matrix:
        X = _as_csr(X)
        given = _check_relations(X, relations)
        if self._feature_size != given:  # predictor.hpp:24-33
            raise ValueError("Told to predict for {} but this->feature_size is {}".format(
                given, self._feature_size))
        return X

    def _device_handles(self):
        """ctypes array of the samples' device handles when every sample lives on the device the
        current engine options name, in their compute dtype; else None (host path)."""
        opts = get_options()
        if not all(isinstance(s, _DeviceFM) and s._on_device_for(opts) for s in self.samples):
            return None
        return (C.c_void_p * len(self.samples))(*[s._h for s in self.samples])

    def _stack(self):
        w0s = np.asarray([s.w0 for s in self.samples], dtype=np.float64)
        ws = np.ascontiguousarray(np.stack([s.w for s in self.samples]), dtype=np.float64)
        Vs = np.ascontiguousarray(np.stack([s.V for s in self.samples]), dtype=np.float64)
        return w0s, ws, Vs

    def _predict_mean(self, X, relations, empty_message: str) -> np.ndarray:
        X = self._check_input(X, relations)
        if not self.samples:
            raise RuntimeError(empty_message)
        ds = _device_dataset(X, list(relations))
        out = np.empty(X.shape[0], dtype=np.float64)
        handles = self._device_handles()
        if handles is not None:
            _lib.check(_lib.lib().myfm_predict_samples_mean(
                ds._h, C.c_int32(int(self._type)), C.c_int32(len(self.samples)), handles, None, C.c_int32(-1),
                _lib.vptr(out)))
            return out
        w0s, ws, Vs = self._stack()
        _lib.check(_lib.lib().myfm_predict_mean(
            ds._h, C.c_int32(int(self._type)), C.c_int32(len(self.samples)), _lib.vptr(w0s),
            _lib.vptr(ws), _lib.vptr(Vs), C.c_int64(self._feature_size), C.c_int32(self._rank),
            _lib.vptr(out)))
        return out

    def predict(self, X, relations: Sequence[RelationBlock] = ()) -> np.ndarray:
        """predictor.hpp:126-147"""
        return self._predict_mean(X, relations, "Empty samples!")

    def predict_parallel(self, X, relations: Sequence[RelationBlock], n_workers: int) -> np.ndarray:
        """predictor.hpp:35-76 — the device loop over samples replaces the thread pool; the
        arithmetic is the same as predict()."""
        return self._predict_mean(X, relations, "Told to predict but no sample available.")

    def predict_parallel_oprobit(self, X, relations: Sequence[RelationBlock], n_workers: int,
                                 cutpoint_index: int) -> np.ndarray:
        """predictor.hpp:78-124"""
        X = self._check_input(X, relations)
        if not self.samples:
            raise RuntimeError("Told to predict but no sample available.")
        if self._type != TaskType.ORDERED:
            raise RuntimeError("predict_parallel_oprobit must be called for oprobit model.")
        cps = np.ascontiguousarray(
            np.stack([s.cutpoints[cutpoint_index] for s in self.samples]), dtype=np.float64)
        ds = _device_dataset(X, list(relations))
        out = np.empty((X.shape[0], cps.shape[1] + 1), dtype=np.float64)
        handles = self._device_handles()
        if handles is not None:
            _lib.check(_lib.lib().myfm_predict_samples_mean(
                ds._h, C.c_int32(int(TaskType.ORDERED)), C.c_int32(len(self.samples)), handles, _lib.vptr(cps),
                C.c_int32(cps.shape[1]), _lib.vptr(out)))
            return out
        w0s, ws, Vs = self._stack()
        _lib.check(_lib.lib().myfm_predict_oprobit_mean(
            ds._h, C.c_int32(len(self.samples)), _lib.vptr(w0s), _lib.vptr(ws), _lib.vptr(Vs),
            _lib.vptr(cps), C.c_int32(cps.shape[1]), C.c_int64(self._feature_size),
            C.c_int32(self._rank), _lib.vptr(out)))
        return out

    def __getstate__(self):
        return (self._rank, self._feature_size, int(self._type), self.samples)

    def __setstate__(self, state) -> None:
        if len(state) != 4:
            raise RuntimeError("invalid state for FMHyperParameters.")
        self.__init__(state[0], state[1], state[2])
        self.samples = list(state[3])


# ------------------------------------------------------------------------------------------------
# LearningHistory — declare_module.hpp:360-377, include/myfm/LearningHistory.hpp
# ------------------------------------------------------------------------------------------------
class LearningHistory:
    def __init__(self) -> None:
        self.hypers: List[FMHyperParameters] = []
        self.train_log_losses: List[float] = []
        self.n_mh_accept: List[int] = []

    def __getstate__(self):
        return (self.hypers, self.train_log_losses, self.n_mh_accept)

    def __setstate__(self, state) -> None:
        if len(state) != 3:
            raise RuntimeError("invalid state for LearningHistory.")
        self.hypers, self.train_log_losses, self.n_mh_accept = (list(s) for s in state)


# ------------------------------------------------------------------------------------------------
# trainer — declare_module.hpp:348-352, :30-45
# ------------------------------------------------------------------------------------------------
class _TrainerHandle:
    """Owns one myfm_trainer_t."""

    def __init__(self, X, relations: Sequence[RelationBlock], y: np.ndarray, random_seed: int,
                 config: FMLearningConfig) -> None:
        self._h = None
        opts = get_options()
        X = _as_csr(X)
        y = np.ascontiguousarray(y, dtype=np.float64)
        csr = _lib.CsrHolder(X)
        rels = _lib.RelationsHolder(relations)
        eo = _lib.EngineOptions()
        eo.dtype, eo.rng, eo.device = _lib.DTYPES[opts.dtype], _lib.RNGS[opts.rng], opts.device
        self.dtype, self.device = opts.dtype, opts.device
        eo.world_size, eo.rank = opts.world_size, opts.rank
        eo.row_offset, eo.n_rows_global = opts.row_offset, opts.n_rows_global or X.shape[0]
        uid = (C.c_char * 128).from_buffer_copy(opts.nccl_unique_id) if opts.nccl_unique_id else None
        eo.nccl_unique_id = C.cast(uid, C.c_void_p) if uid is not None else None
        levels = None
        if opts.column_level is not None:
            levels = np.ascontiguousarray(opts.column_level, dtype=np.int32)
            eo.column_level, eo.n_column_level = _lib.ptr(levels, C.c_int32), levels.shape[0]
        row_ids = None
        if opts.row_ids is not None:
            row_ids = np.ascontiguousarray(opts.row_ids, dtype=np.int64)
            if row_ids.shape[0] != X.shape[0]:
                raise ValueError("row_ids must have one entry per training row of this shard.")
            eo.row_ids, eo.n_row_ids = _lib.ptr(row_ids, C.c_int64), row_ids.shape[0]
        cfg_struct = config._as_struct()
        h = C.c_void_p()
        _lib.check(_lib.lib().myfm_trainer_create(
            C.byref(h), C.byref(csr.struct), C.c_int32(rels.n), rels.array, _lib.vptr(y),
            C.c_int64(y.shape[0]), C.c_int32(int(random_seed)), C.byref(cfg_struct), C.byref(eo)))
        self._h = h
        self.config = config
        self.task_type = TaskType(int(config.task_type))
        self.n_cutpoint_groups = len(config.cutpoint_groups)
        self.cutpoint_sizes = [int(g[0]) - 1 for g in config.cutpoint_groups]
        self.rank = -1

    def __del__(self) -> None:
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.lib().myfm_trainer_destroy(self._h)
            self._h = None

    def init_fm(self, rank: int, init_std: float) -> None:
        _lib.check(_lib.lib().myfm_trainer_init_fm(self._h, int(rank), float(init_std)))
        n, d, k, g = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().myfm_trainer_dims(self._h, C.byref(n), C.byref(d), C.byref(k), C.byref(g)))
        self.n_train, self.dim_all, self.rank, self.n_groups = n.value, d.value, k.value, g.value

    def step(self, n: int = 1) -> None:
        _lib.check(_lib.lib().myfm_trainer_step(self._h, int(n)))

    def sync(self) -> None:
        _lib.check(_lib.lib().myfm_trainer_sync(self._h))

    def timed_steps(self, n: int) -> float:
        """n sweeps bracketed by CUDA events on the trainer's stream; device milliseconds."""
        ms = C.c_double()
        _lib.check(_lib.lib().myfm_trainer_timed_steps(self._h, C.c_int32(int(n)), C.byref(ms)))
        return float(ms.value)

    def get_fm(self) -> Tuple[float, np.ndarray, np.ndarray, List[np.ndarray]]:
        w0 = C.c_double()
        w = np.empty(self.dim_all, dtype=np.float64)
        V = np.empty((self.dim_all, self.rank), dtype=np.float64)
        _lib.check(_lib.lib().myfm_trainer_get_fm(self._h, C.byref(w0), _lib.vptr(w), _lib.vptr(V)))
        cps = []
        if self.task_type == TaskType.ORDERED:
            for g, size in enumerate(self.cutpoint_sizes):
                a = np.empty(size, dtype=np.float64)
                _lib.check(_lib.lib().myfm_trainer_get_cutpoints(self._h, C.c_int32(g), _lib.vptr(a)))
                cps.append(a)
        return w0.value, w, V, cps

    def snapshot(self) -> "_DeviceFM":
        """The current sample, copied device to device (the kept-sample copy of FMTrainer.hpp:74-76)."""
        h = C.c_void_p()
        _lib.check(_lib.lib().myfm_trainer_snapshot(self._h, C.byref(h)))
        cps = []
        if self.task_type == TaskType.ORDERED:
            for g, size in enumerate(self.cutpoint_sizes):
                a = np.empty(size, dtype=np.float64)
                _lib.check(_lib.lib().myfm_trainer_get_cutpoints(self._h, C.c_int32(g), _lib.vptr(a)))
                cps.append(a)
        return _DeviceFM(h.value, self.get_w0(), cps, self.dim_all, self.rank, self.dtype, self.device)

    def get_w0(self) -> float:
        w0 = C.c_double()
        _lib.check(_lib.lib().myfm_trainer_get_fm(self._h, C.byref(w0), None, None))
        return w0.value

    def get_hyper(self) -> FMHyperParameters:
        G, K = self.n_groups, self.rank
        alpha = C.c_double()
        mu_w, lambda_w = np.empty(G), np.empty(G)
        mu_V, lambda_V = np.empty((G, K)), np.empty((G, K))
        _lib.check(_lib.lib().myfm_trainer_get_hyper(
            self._h, C.byref(alpha), _lib.vptr(mu_w), _lib.vptr(lambda_w), _lib.vptr(mu_V),
            _lib.vptr(lambda_V)))
        return FMHyperParameters(alpha.value, mu_w, lambda_w, mu_V, lambda_V)

    def get_e(self) -> np.ndarray:
        e = np.empty(self.n_train, dtype=np.float64)
        _lib.check(_lib.lib().myfm_trainer_get_e(self._h, _lib.vptr(e)))
        return e

    def get_q(self) -> np.ndarray:
        q = np.empty(self.n_train, dtype=np.float64)
        _lib.check(_lib.lib().myfm_trainer_get_q(self._h, _lib.vptr(q)))
        return q

    def set_state(self, w0=None, w=None, V=None, hyper: Optional["FMHyperParameters"] = None,
                  e=None) -> None:
        """Overwrite parts of the chain state (warm start / teacher-forced parity tests)."""
        keep = []

        def arr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return _lib.vptr(a)

        alpha = arr(np.asarray([hyper.alpha])) if hyper is not None else None
        _lib.check(_lib.lib().myfm_trainer_set_state(
            self._h, arr(np.asarray([w0])) if w0 is not None else None, arr(w), arr(V), alpha,
            arr(hyper.mu_w) if hyper is not None else None,
            arr(hyper.lambda_w) if hyper is not None else None,
            arr(hyper.mu_V) if hyper is not None else None,
            arr(hyper.lambda_V) if hyper is not None else None, arr(e)))

    def get_variates(self) -> np.ndarray:
        """Standardised variates the most recent sweep consumed (reference draw order)."""
        n = C.c_int64()
        _lib.check(_lib.lib().myfm_trainer_get_variates(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.empty(n.value, dtype=np.float64)
        _lib.check(_lib.lib().myfm_trainer_get_variates(self._h, _lib.vptr(out), C.c_int64(n.value), C.byref(n)))
        return out

    def mh_accept(self, g: int) -> int:
        n = C.c_int64()
        _lib.check(_lib.lib().myfm_trainer_mh_accept(self._h, C.c_int32(g), C.byref(n)))
        return int(n.value)

    def launch_count(self) -> int:
        n = C.c_int64()
        _lib.check(_lib.lib().myfm_trainer_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def sweep_path(self) -> int:
        """0 = general dependency-level kernels, 1 = field path (csrc/field_sweep.cuh)."""
        n = C.c_int32()
        _lib.check(_lib.lib().myfm_trainer_sweep_path(self._h, C.byref(n)))
        return int(n.value)

    def set_profiling(self, on: bool) -> None:
        _lib.check(_lib.lib().myfm_trainer_set_profiling(self._h, C.c_int32(int(on))))

    def kernel_ms(self, family: int) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        _lib.check(_lib.lib().myfm_trainer_kernel_ms(self._h, C.c_int32(family), C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)


class FMTrainer:
    """declare_module.hpp:348-352 — (X, relations, y, random_seed, config)."""

    def __init__(self, X, relations: Sequence[RelationBlock], y: np.ndarray, random_seed: int,
                 config: FMLearningConfig) -> None:
        self._handle = _TrainerHandle(X, relations, y, random_seed, config)

    def create_FM(self, rank: int, init_std: float) -> FM:
        self._handle.init_fm(rank, init_std)
        return _LiveFM(self._handle).freeze()

    def create_Hyper(self, rank: int) -> FMHyperParameters:
        G = self._handle.config.n_groups
        return FMHyperParameters(0.0, np.zeros(G), np.zeros(G), np.zeros((G, rank)), np.zeros((G, rank)))


def create_train_fm(
    rank: int,
    init_std: float,
    X,
    relations: Sequence[RelationBlock],
    y: np.ndarray,
    random_seed: int,
    config: FMLearningConfig,
    callback: Callable[[int, FM, FMHyperParameters, LearningHistory], bool],
) -> Tuple[Predictor, LearningHistory]:
    """create and train fm.  (declare_module.hpp:30-45 -> FMTrainer.hpp:56-87)"""
    trainer = _TrainerHandle(X, relations, y, random_seed, config)
    trainer.init_fm(rank, init_std)
    predictor = Predictor(rank, trainer.dim_all, int(config.task_type))
    history = LearningHistory()
    n_iter, n_kept = int(config.n_iter), int(config.n_kept_samples)
    # A callback that only observes `hyper` / `history` (attribute `observer`: True, or a predicate of the
    # iteration) lets the next sweep start on the device before the callback's host work: `fm` is then NOT
    # valid inside it.  MyFM*.fit() marks its own progress-bar callback on the iterations it does not report.
    observer = getattr(callback, "observer", False)
    if os.environ.get("MYFM_B200_NO_RUN_AHEAD", os.environ.get("MYFM_NO_RUN_AHEAD", "0")) == "1":
        observer = False
    in_flight = False
    for it in range(n_iter):
        if not in_flight:
            trainer.step(1)
        in_flight = False
        live = _LiveFM(trainer)
        if n_iter <= it + n_kept:
            predictor.samples.append(trainer.snapshot())
        hyper = trainer.get_hyper()
        history.hypers.append(hyper)
        if it + 1 < n_iter and (observer(it) if callable(observer) else observer):
            trainer.step(1)
            in_flight = True
        if callback(it, live, hyper, history):
            break
    trainer.sync()  # raises if the device flagged an error (every get_hyper() above checks as well)
    for g in range(trainer.n_cutpoint_groups):
        history.n_mh_accept.append(trainer.mh_accept(g))
    return predictor, history


def create_train_vfm(*args, **kwargs):
    """Variational inference (reference include/myfm/variational.hpp) is a different algorithm
    and outside the accelerated hot path (SURVEY.md §8 f#4)."""
    raise NotImplementedError("the variational trainer is not part of myfm_b200")


# ------------------------------------------------------------------------------------------------
# include/myfm/util.hpp:80-115
# ------------------------------------------------------------------------------------------------
_SQRT2 = 1.4142135623730951
_SQRT2PI = 1.4142135623730951 * 1.7724538509055159


def mean_var_truncated_normal_left(mu: float) -> Tuple[float, float, float]:
    mu = float(mu)
    mu_square = mu * mu / 2
    if mu > 0:
        Z = 1 - special.erf(-mu / _SQRT2)
        phi_Z = 2 * math.exp(-mu_square) / _SQRT2PI / Z
        lnZ = math.log(Z)
    else:
        Z = float(special.erfcx(-mu / _SQRT2))
        phi_Z = 2 / Z / _SQRT2PI
        lnZ = math.log(Z) - mu_square
    return (mu + phi_Z, 1 - mu * phi_Z - phi_Z * phi_Z, lnZ)


def mean_var_truncated_normal_right(mu: float) -> Tuple[float, float, float]:
    mean, var, lnZ = mean_var_truncated_normal_left(-mu)
    return (-mean, var, lnZ)
