"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes driver for ``oracle/_build/liboracle.so`` (the Eigen-free CPU restatement of the
reference sampler, ``oracle/fm_oracle.hpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this module; nothing under ``myfm_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy import sparse as sps

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

TASKS = {"regression": 0, "classification": 1, "ordered": 2}
DTYPES = {"f32": 0, "f64": 1, "float32": 0, "float64": 1}


class _RelationDesc(C.Structure):
    _fields_ = [
        ("original_to_block", C.POINTER(C.c_int64)),
        ("mapper_size", C.c_int64),
        ("block_size", C.c_int64),
        ("feature_size", C.c_int64),
        ("indptr", C.POINTER(C.c_int64)),
        ("indices", C.POINTER(C.c_int32)),
        ("data", C.POINTER(C.c_double)),
    ]


class _ConfigDesc(C.Structure):
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


def build(force: bool = False) -> str:
    """Compile the oracle (and, when /root/reference is present, oracle/_ref/faddeeva.o)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("fm_oracle.hpp", "oracle_capi.cpp")
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_chain_create.restype = C.c_void_p
        L.oracle_chain_timed_steps.restype = C.c_double
        L.oracle_chain_dim_all.restype = C.c_int64
        L.oracle_chain_n_train.restype = C.c_int64
        L.oracle_chain_mh_accept.restype = C.c_int64
        L.oracle_erfcx.restype = C.c_double
        L.oracle_erfcx.argtypes = [C.c_double]
        _lib = L
    return _lib


def _raise_last() -> None:
    msg = lib().oracle_last_error().decode()
    if msg.startswith("invalid_argument: "):
        raise ValueError(msg[len("invalid_argument: "):])
    raise RuntimeError(msg.split(": ", 1)[-1])


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class _CsrArgs:
    """Keeps the converted arrays of one CSR matrix alive while C reads them."""

    def __init__(self, X) -> None:
        X = sps.csr_matrix(X)
        self.shape = X.shape
        self.indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(X.indices, dtype=np.int32)
        self.data = np.ascontiguousarray(X.data, dtype=np.float64)

    def args(self):
        return (
            C.c_int64(self.shape[0]),
            C.c_int64(self.shape[1]),
            _ptr(self.indptr, C.c_int64),
            _ptr(self.indices, C.c_int32),
            _ptr(self.data, C.c_double),
        )


def _split_relation(rel) -> Tuple[np.ndarray, sps.csr_matrix]:
    if isinstance(rel, tuple):
        return np.asarray(rel[0]), sps.csr_matrix(rel[1])
    return np.asarray(rel.original_to_block), sps.csr_matrix(rel.data)


class _RelArgs:
    def __init__(self, X_rel: Sequence) -> None:
        self.keep = []
        self.array = (_RelationDesc * max(1, len(X_rel)))()
        for b, rel in enumerate(X_rel):
            omap, mat = _split_relation(rel)
            omap = np.ascontiguousarray(omap, dtype=np.int64)
            csr = _CsrArgs(mat)
            self.keep.append((omap, csr))
            d = self.array[b]
            d.original_to_block = _ptr(omap, C.c_int64)
            d.mapper_size = omap.shape[0]
            d.block_size, d.feature_size = csr.shape
            d.indptr = _ptr(csr.indptr, C.c_int64)
            d.indices = _ptr(csr.indices, C.c_int32)
            d.data = _ptr(csr.data, C.c_double)
        self.n = len(X_rel)


class OracleChain:
    """One Gibbs chain of the restatement: construct == create_train_fm up to the first
    iteration (trainer ctor, create_FM, create_Hyper, initialize_hyper, initialize_e);
    ``step()`` == one ``update_all`` (reference include/myfm/BaseFMTrainer.hpp:135-152)."""

    def __init__(
        self,
        X,
        y: np.ndarray,
        rank: int,
        *,
        X_rel: Sequence = (),
        dtype: str = "f64",
        task: str = "regression",
        seed: int = 42,
        init_std: float = 0.1,
        group_index: Optional[Sequence[int]] = None,
        group_shapes: Optional[Sequence[int]] = None,
        n_iter: int = 100,
        n_kept_samples: Optional[int] = None,
        alpha_0: float = 1.0,
        beta_0: float = 1.0,
        gamma_0: float = 1.0,
        mu_0: float = 0.0,
        reg_0: float = 1.0,
        fit_w0: bool = True,
        fit_linear: bool = True,
        nu_oprobit: float = 5,
        cutpoint_groups: Optional[List[Tuple[int, Sequence[int]]]] = None,
    ) -> None:
        self._h = None
        L = lib()
        self.dtype = dtype
        Xa = _CsrArgs(X)
        rels = _RelArgs(list(X_rel))
        dim_all = Xa.shape[1] + sum(k[1].shape[1] for k in rels.keep)
        if group_index is None:
            if group_shapes is not None:
                group_index = [g for g, n in enumerate(group_shapes) for _ in range(n)]
            else:
                group_index = [0] * dim_all
        gi = np.ascontiguousarray(group_index, dtype=np.int64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if task == "ordered" and cutpoint_groups is None:
            cutpoint_groups = [(int(y.max()) + 1, np.arange(y.shape[0]))]
        cutpoint_groups = cutpoint_groups or []
        if n_kept_samples is None:
            n_kept_samples = min(max(n_iter - 5, 5), n_iter)
        cfg = _ConfigDesc()
        cfg.alpha_0, cfg.beta_0, cfg.gamma_0, cfg.mu_0, cfg.reg_0 = alpha_0, beta_0, gamma_0, mu_0, reg_0
        cfg.task_type = TASKS[task]
        cfg.nu_oprobit = nu_oprobit
        cfg.fit_w0, cfg.fit_linear = int(fit_w0), int(fit_linear)
        cfg.n_iter, cfg.n_kept_samples = n_iter, n_kept_samples
        cfg.cutpoint_scale = 10.0
        cfg.group_index, cfg.n_group_index = _ptr(gi, C.c_int64), gi.shape[0]
        n_cg = len(cutpoint_groups)
        cg_ncls = np.asarray([c[0] for c in cutpoint_groups] or [0], dtype=np.int32)
        cg_rows = [np.ascontiguousarray(c[1], dtype=np.int64) for c in cutpoint_groups]
        cg_ptrs = (C.POINTER(C.c_int64) * max(1, n_cg))(*[_ptr(r, C.c_int64) for r in cg_rows])
        cg_lens = np.asarray([r.shape[0] for r in cg_rows] or [0], dtype=np.int64)
        cfg.n_cutpoint_groups = n_cg
        cfg.cutpoint_n_class = _ptr(cg_ncls, C.c_int32)
        cfg.cutpoint_index = C.cast(cg_ptrs, C.POINTER(C.POINTER(C.c_int64)))
        cfg.cutpoint_index_len = _ptr(cg_lens, C.c_int64)

        h = L.oracle_chain_create(
            C.c_int(DTYPES[dtype]), C.c_int(rank), C.c_double(init_std), *Xa.args(),
            C.c_int(rels.n), rels.array, _ptr(y, C.c_double), C.c_int64(y.shape[0]),
            C.c_int(seed), C.byref(cfg),
        )
        if not h:
            _raise_last()
        self._h = C.c_void_p(h)
        self.n_iter, self.n_kept_samples = n_iter, n_kept_samples
        self.dim_all = int(L.oracle_chain_dim_all(self._h))
        self.rank = int(L.oracle_chain_rank(self._h))
        self.n_groups = int(L.oracle_chain_n_groups(self._h))
        self.n_train = int(L.oracle_chain_n_train(self._h))

    def __del__(self) -> None:
        if getattr(self, "_h", None):
            lib().oracle_chain_destroy(self._h)
            self._h = None

    def step(self) -> None:
        if lib().oracle_chain_step(self._h) != 0:
            _raise_last()

    def timed_steps(self, n: int) -> float:
        t = lib().oracle_chain_timed_steps(self._h, C.c_int(n))
        if t < 0:
            _raise_last()
        return float(t)

    def fm(self) -> Tuple[float, np.ndarray, np.ndarray]:
        w0 = C.c_double()
        w = np.empty(self.dim_all)
        V = np.empty((self.rank, self.dim_all))  # column-major (dim_all x rank) == C-order (rank x dim_all)
        lib().oracle_chain_get_fm(self._h, C.byref(w0), _ptr(w, C.c_double), _ptr(V, C.c_double))
        return w0.value, w, np.ascontiguousarray(V.T)

    def cutpoints(self) -> List[np.ndarray]:
        L = lib()
        out = []
        for g in range(L.oracle_chain_n_cutpoint_groups(self._h)):
            a = np.empty(L.oracle_chain_cutpoint_len(self._h, C.c_int(g)))
            L.oracle_chain_get_cutpoints(self._h, C.c_int(g), _ptr(a, C.c_double))
            out.append(a)
        return out

    def hyper(self) -> Dict[str, np.ndarray]:
        alpha = C.c_double()
        G, K = self.n_groups, self.rank
        mu_w, lambda_w = np.empty(G), np.empty(G)
        mu_V, lambda_V = np.empty((K, G)), np.empty((K, G))
        lib().oracle_chain_get_hyper(
            self._h, C.byref(alpha), _ptr(mu_w, C.c_double), _ptr(lambda_w, C.c_double),
            _ptr(mu_V, C.c_double), _ptr(lambda_V, C.c_double),
        )
        return dict(alpha=alpha.value, mu_w=mu_w, lambda_w=lambda_w,
                    mu_V=np.ascontiguousarray(mu_V.T), lambda_V=np.ascontiguousarray(lambda_V.T))

    def e(self) -> np.ndarray:
        a = np.empty(self.n_train)
        lib().oracle_chain_get_e(self._h, _ptr(a, C.c_double))
        return a

    def q(self) -> np.ndarray:
        a = np.empty(self.n_train)
        lib().oracle_chain_get_q(self._h, _ptr(a, C.c_double))
        return a

    def mh_accept(self, g: int = 0) -> int:
        return int(lib().oracle_chain_mh_accept(self._h, C.c_int(g)))

    def run(self, n_iter: Optional[int] = None):
        """learn_with_callback without a callback (reference FMTrainer.hpp:56-87): returns
        (kept samples, per-iteration hypers); a sample is (w0, w, V, cutpoints)."""
        n_iter = self.n_iter if n_iter is None else n_iter
        samples, hypers = [], []
        for it in range(n_iter):
            self.step()
            if n_iter <= it + self.n_kept_samples:
                samples.append(self.fm() + (self.cutpoints(),))
            hypers.append(self.hyper())
        return samples, hypers


def predict_score(dtype: str, w0: float, w: np.ndarray, V: np.ndarray, X, X_rel: Sequence = ()) -> np.ndarray:
    """FM.predict_score of the restatement (reference include/myfm/FM.hpp:54-136)."""
    Xa = _CsrArgs(X)
    rels = _RelArgs(list(X_rel))
    w = np.ascontiguousarray(w, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64).reshape(w.shape[0], -1)
    Vt = np.ascontiguousarray(V.T)  # column-major (dim_all x rank)
    out = np.empty(Xa.shape[0])
    rc = lib().oracle_predict_score(
        C.c_int(DTYPES[dtype]), C.c_double(w0), _ptr(w, C.c_double), _ptr(Vt, C.c_double),
        C.c_int64(w.shape[0]), C.c_int(V.shape[1]), *Xa.args(), C.c_int(rels.n), rels.array,
        _ptr(out, C.c_double),
    )
    if rc != 0:
        _raise_last()
    return out


def kat_normal(dtype: str, seed: int, n: int) -> np.ndarray:
    out = np.empty(n)
    lib().oracle_kat_normal(C.c_int(DTYPES[dtype]), C.c_int(seed), C.c_int(n), _ptr(out, C.c_double))
    return out


def tn_draws(dtype: str, seed: int, kind: int, a: float, b: float, n: int) -> np.ndarray:
    out = np.empty(n)
    lib().oracle_tn_draws(C.c_int(DTYPES[dtype]), C.c_int(seed), C.c_int(kind), C.c_double(a),
                          C.c_double(b), C.c_int(n), _ptr(out, C.c_double))
    return out


def erfcx(x: float) -> float:
    return float(lib().oracle_erfcx(C.c_double(x)))
