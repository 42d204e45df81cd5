"""ctypes binding of the C ABI declared in ``include/myfm_b200.h``.

The shared object is built in-tree by ``myfm_b200/csrc/build.py`` (``__graft_entry__.build()``).
There is no CPU fallback: if the library is missing the import of the compute path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse as sps

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmyfm_b200.so")

OK, ERR_INVALID_ARGUMENT, ERR_RUNTIME, ERR_CUDA = 0, 1, 2, 3
DTYPE_F32, DTYPE_F64 = 0, 1
RNG_MT19937, RNG_PHILOX = 0, 1
DTYPES = {"f32": DTYPE_F32, "float32": DTYPE_F32, "f64": DTYPE_F64, "float64": DTYPE_F64}
RNGS = {"mt19937": RNG_MT19937, "philox": RNG_PHILOX}


class CudaEngineError(RuntimeError):
    """The CUDA engine could not run (library missing, no device, CUDA failure)."""


class Csr(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64),
        ("n_cols", C.c_int64),
        ("indptr", C.POINTER(C.c_int64)),
        ("indices", C.POINTER(C.c_int32)),
        ("data", C.POINTER(C.c_double)),
    ]


class Relation(C.Structure):
    _fields_ = [
        ("original_to_block", C.POINTER(C.c_int64)),
        ("mapper_size", C.c_int64),
        ("block", Csr),
    ]


class Config(C.Structure):
    _fields_ = [
        ("alpha_0", C.c_double),
        ("beta_0", C.c_double),
        ("gamma_0", C.c_double),
        ("mu_0", C.c_double),
        ("reg_0", C.c_double),
        ("task_type", C.c_int32),
        ("nu_oprobit", C.c_double),
        ("fit_w0", C.c_int32),
        ("fit_linear", C.c_int32),
        ("n_iter", C.c_int32),
        ("n_kept_samples", C.c_int32),
        ("cutpoint_scale", C.c_double),
        ("group_index", C.POINTER(C.c_int64)),
        ("n_group_index", C.c_int64),
        ("n_cutpoint_groups", C.c_int32),
        ("cutpoint_n_class", C.POINTER(C.c_int32)),
        ("cutpoint_index", C.POINTER(C.POINTER(C.c_int64))),
        ("cutpoint_index_len", C.POINTER(C.c_int64)),
    ]


class EngineOptions(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("rng", C.c_int32),
        ("device", C.c_int32),
        ("world_size", C.c_int32),
        ("rank", C.c_int32),
        ("row_offset", C.c_int64),
        ("n_rows_global", C.c_int64),
        ("nccl_unique_id", C.c_void_p),
        ("column_level", C.POINTER(C.c_int32)),
        ("n_column_level", C.c_int64),
        ("row_ids", C.POINTER(C.c_int64)),
        ("n_row_ids", C.c_int64),
    ]


# every symbol include/myfm_b200.h declares (tests/test_cabi.py checks the two stay in sync)
EXPORTS = [
    "myfm_last_error", "myfm_device_count", "myfm_config_validate",
    "myfm_trainer_create", "myfm_trainer_destroy", "myfm_trainer_init_fm", "myfm_trainer_step",
    "myfm_trainer_sync", "myfm_trainer_timed_steps", "myfm_trainer_dims", "myfm_trainer_get_fm", "myfm_trainer_get_cutpoints",
    "myfm_trainer_get_hyper", "myfm_trainer_get_e", "myfm_trainer_get_q", "myfm_trainer_set_state",
    "myfm_trainer_mh_accept", "myfm_trainer_get_variates",
    "myfm_trainer_sweep_path", "myfm_trainer_launch_count", "myfm_trainer_kernel_ms", "myfm_trainer_set_profiling",
    "myfm_dataset_create", "myfm_dataset_destroy", "myfm_predict_score", "myfm_predict_mean",
    "myfm_predict_oprobit_mean", "myfm_trainer_predict_score", "myfm_rng_fill",
    "myfm_trainer_snapshot", "myfm_sample_destroy", "myfm_sample_get", "myfm_predict_samples_mean",
    "myfm_level_schedule", "myfm_level_relax", "myfm_host_transpose", "myfm_set_host_threads", "myfm_nccl_unique_id",
    "myfm_mt_jump_taps",
    "myfm_evaluator_create", "myfm_evaluator_destroy", "myfm_evaluator_step", "myfm_evaluator_get_sums",
]

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CudaEngineError(
                f"{LIB_PATH} is missing: the CUDA engine has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "myfm_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        L.myfm_last_error.restype = C.c_char_p
        L.myfm_trainer_destroy.restype = None
        L.myfm_dataset_destroy.restype = None
        L.myfm_sample_destroy.restype = None
        L.myfm_evaluator_destroy.restype = None
        L.myfm_evaluator_destroy.argtypes = [C.c_void_p]
        L.myfm_evaluator_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                            C.c_double, C.c_double, C.c_double]
        L.myfm_evaluator_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.myfm_evaluator_get_sums.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.myfm_sample_destroy.argtypes = [C.c_void_p]
        L.myfm_predict_score.argtypes = [
            C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.myfm_predict_mean.argtypes = [
            C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
            C.c_int32, C.c_void_p]
        L.myfm_predict_oprobit_mean.argtypes = [
            C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
            C.c_int64, C.c_int32, C.c_void_p]
        L.myfm_trainer_init_fm.argtypes = [C.c_void_p, C.c_int32, C.c_double]
        L.myfm_trainer_step.argtypes = [C.c_void_p, C.c_int32]
        for name in ("myfm_trainer_sync", "myfm_trainer_destroy", "myfm_dataset_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Maps the C error codes onto the exceptions pybind11 raises for the reference."""
    if rc == OK:
        return
    msg = lib().myfm_last_error().decode(errors="replace")
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if rc == ERR_CUDA:
        raise CudaEngineError(msg)
    raise RuntimeError(msg)


def device_count() -> int:
    n = C.c_int32(0)
    check(lib().myfm_device_count(C.byref(n)))
    return int(n.value)


def ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def vptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class CsrHolder:
    """A scipy matrix converted to the ABI's layout; keeps the arrays alive."""

    def __init__(self, X) -> None:
        if not sps.isspmatrix_csr(X):
            X = sps.csr_matrix(X)
        if not X.has_canonical_format:
            # the same (row, col) listed twice would make the reference's serial column pass
            # read its own partial update; sum them, which is what the model means
            X = X.copy()
            X.sum_duplicates()
        self.shape = X.shape
        self.indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(X.indices, dtype=np.int32)
        self.data = np.ascontiguousarray(X.data, dtype=np.float64)
        self.struct = Csr(
            self.shape[0], self.shape[1], ptr(self.indptr, C.c_int64),
            ptr(self.indices, C.c_int32), ptr(self.data, C.c_double),
        )


class RelationsHolder:
    def __init__(self, relations: Sequence) -> None:
        self.keep: List[Tuple[np.ndarray, CsrHolder]] = []
        self.array = (Relation * max(1, len(relations)))()
        self.n = len(relations)
        for b, rel in enumerate(relations):
            omap = np.ascontiguousarray(rel._map, dtype=np.int64)
            csr = CsrHolder(rel._data)
            self.keep.append((omap, csr))
            self.array[b].original_to_block = ptr(omap, C.c_int64)
            self.array[b].mapper_size = omap.shape[0]
            self.array[b].block = csr.struct
