"""The C-ABI library loads without a GPU and exports every symbol include/myfm_b200.h declares;
host-only entry points (config validation, level schedule, RNG stream) work on the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sps

from helpers import middle_data, movielens_like, toy_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "myfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(myfm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from myfm_b200 import _lib

    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/myfm_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_device_probe_never_fails():
    from myfm_b200 import _lib

    assert _lib.device_count() >= 0


def test_no_cpu_fallback_when_no_device():
    import myfm_b200
    from myfm_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    X, y = toy_matrix()
    with pytest.raises(_lib.CudaEngineError, match="no CPU fallback"):
        myfm_b200.MyFMRegressor(2).fit(X, y, n_iter=2)
    from myfm_b200._myfm import FM

    with pytest.raises(_lib.CudaEngineError):
        FM(0.0, np.zeros(9), np.zeros((9, 2))).predict_score(X, [])


def level_schedule(X):
    from myfm_b200 import _lib

    h = _lib.CsrHolder(X)
    level = np.empty(X.shape[1], dtype=np.int32)
    n = C.c_int32()
    _lib.check(_lib.lib().myfm_level_schedule(C.byref(h.struct), _lib.vptr(level), C.byref(n)))
    return level, n.value


def test_level_schedule_one_hot_fields():
    X, _, shapes = movielens_like(2000, 50, 20, 2, seed=0)
    level, n = level_schedule(X)
    assert n == 2
    assert np.all(level[: shapes[0]] == 0) and np.all(level[shapes[0]:] == 1)


def test_level_schedule_respects_serial_order():
    """Columns of one level never share a row, and a column's level exceeds that of every earlier
    column it shares a row with (so level order == the reference's index order)."""
    rng = np.random.default_rng(0)
    X = sps.random(60, 25, 0.15, format="csr", random_state=1)
    level, n = level_schedule(X)
    dense = X.toarray() != 0
    for j in range(X.shape[1]):
        for i in range(j):
            if np.any(dense[:, i] & dense[:, j]):
                assert level[i] < level[j]
    for lv in range(n):
        cols = np.where(level == lv)[0]
        assert np.all(dense[:, cols].sum(axis=1) <= 1)
    Xd, _ = toy_matrix()
    level, n = level_schedule(Xd)
    assert n == 3 and level[0] == 0 and set(level[1:5]) == {1} and set(level[5:]) == {2}
    Xm, _ = middle_data(50)
    assert level_schedule(Xm)[1] == 3


def rng_fill(dtype, seed, n_skip, kinds, shapes):
    from myfm_b200 import _lib

    kinds = np.ascontiguousarray(kinds, dtype=np.int32)
    shapes = np.ascontiguousarray(shapes, dtype=np.float64)
    out = np.empty(kinds.shape[0])
    _lib.check(_lib.lib().myfm_rng_fill(C.c_int32(_lib.DTYPES[dtype]), C.c_int32(seed), C.c_int64(n_skip),
                                        _lib.vptr(kinds), _lib.vptr(shapes), C.c_int64(kinds.shape[0]),
                                        _lib.vptr(out)))
    return out


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_engine_rng_stream_is_the_libstdcxx_stream(oracle, dtype):
    """A fresh normal_distribution per draw drops the second polar variate, so the engine's k-th
    fresh normal is element 2k of one persistent distribution on the same mt19937."""
    persistent = oracle.kat_normal(dtype, 42, 64)
    fresh = rng_fill(dtype, 42, 0, [0] * 32, [0] * 32)
    np.testing.assert_array_equal(fresh, persistent[0::2])
    # after an even number of persistent draws the stream is aligned the same way
    fresh_after = rng_fill(dtype, 42, 10, [0] * 8, [0] * 8)
    np.testing.assert_array_equal(fresh_after, persistent[10:26:2])


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_engine_gamma_is_standardised(dtype):
    g = rng_fill(dtype, 7, 0, [1] * 2000, [2.5] * 2000)
    assert abs(g.mean() - 2.5) < 0.15 and abs(g.var() - 2.5) < 0.5
    big = rng_fill(dtype, 7, 0, [1] * 200, [5.0e6] * 200)
    assert abs(big.mean() / 5.0e6 - 1) < 1e-3


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_host_transpose_threads(threads):
    """The multi-threaded counting sort of the setup (csrc/host_data.hpp) against scipy: columns
    list their rows in ascending order whatever the number of threads (BaseFMTrainer.hpp:61)."""
    rng = np.random.default_rng(threads)
    X = sps.random(5000, 300, 0.02, format="csr", random_state=int(rng.integers(1 << 30)))
    X.sort_indices()
    from myfm_b200 import _lib

    h = _lib.CsrHolder(X)
    indptr = np.empty(X.shape[1] + 1, dtype=np.int64)
    indices = np.empty(X.nnz, dtype=np.int32)
    data = np.empty(X.nnz, dtype=np.float64)
    _lib.check(_lib.lib().myfm_set_host_threads(C.c_int32(threads)))
    try:
        _lib.check(_lib.lib().myfm_host_transpose(C.byref(h.struct), _lib.vptr(indptr), _lib.vptr(indices),
                                                  _lib.vptr(data)))
    finally:
        _lib.check(_lib.lib().myfm_set_host_threads(C.c_int32(0)))
    ref = X.tocsc()
    ref.sort_indices()
    np.testing.assert_array_equal(indptr, ref.indptr)
    np.testing.assert_array_equal(indices, ref.indices)
    np.testing.assert_array_equal(data, ref.data)


@pytest.mark.parametrize("threads", [1, 5])
def test_row_parallel_level_schedule_equals_serial(threads):
    """compute_levels_by_rows (row side, several threads, what the trainer uses) against the serial
    column-order recurrence (myfm_level_relax from all-zero lower bounds), on shapes with long
    dependency chains; rows with unsorted columns take the serial route and give the same answer."""
    from myfm_b200 import _lib

    rng = np.random.default_rng(3)
    cases = [sps.random(3000, 120, 0.02, format="csr", random_state=1),
             sps.random(500, 40, 0.3, format="csr", random_state=2),
             movielens_like(4000, 90, 35, 2, seed=4)[0],
             sps.csr_matrix(np.tril(np.ones((30, 30))))]
    _lib.check(_lib.lib().myfm_set_host_threads(C.c_int32(threads)))
    try:
        for X in cases:
            X = sps.csr_matrix(X)
            X.sort_indices()
            got, n = level_schedule(X)
            h = _lib.CsrHolder(X)
            serial = np.zeros(X.shape[1], dtype=np.int32)
            nl, changed = C.c_int32(), C.c_int32()
            _lib.check(_lib.lib().myfm_level_relax(C.byref(h.struct), _lib.vptr(serial), C.byref(nl), C.byref(changed)))
            np.testing.assert_array_equal(got, serial)
            assert n == nl.value
            # the same matrix with every row's entries reversed (unsorted): serial fallback, same levels
            Xr = X.copy()
            for r in range(Xr.shape[0]):
                lo, hi = Xr.indptr[r], Xr.indptr[r + 1]
                Xr.indices[lo:hi] = Xr.indices[lo:hi][::-1].copy()
                Xr.data[lo:hi] = Xr.data[lo:hi][::-1].copy()
            Xr.has_sorted_indices = False
            got_r, _ = level_schedule(Xr)
            np.testing.assert_array_equal(got_r, serial)
    finally:
        _lib.check(_lib.lib().myfm_set_host_threads(C.c_int32(0)))


@pytest.mark.parametrize("n", [20561, 31 * 32915 + 20560, (1 << 33) + 12345])
def test_mt19937_jump_polynomial(n):
    """csrc/mt_jump.hpp: g = t^n mod phi(MT19937).  The generator's output words satisfy
    out[a + n + j] = XOR over the taps i of g of out[a + i + j]; checked against numpy's MT19937
    (the same generator as std::mt19937: identical tempered 32-bit outputs).  The device generator's
    lanes (mt_device.cuh: k_mt_farm) reach their next segment with exactly this identity."""
    from myfm_b200 import _lib

    count = C.c_int32()
    _lib.check(_lib.lib().myfm_mt_jump_taps(C.c_uint64(n), None, C.c_int32(0), C.byref(count)))
    taps = np.zeros(count.value, dtype=np.uint16)
    _lib.check(_lib.lib().myfm_mt_jump_taps(C.c_uint64(n), _lib.vptr(taps), C.c_int32(taps.shape[0]), C.byref(count)))
    assert 9000 < taps.shape[0] < 11000 or n < 30000  # about half of the 19937 coefficients
    assert np.all(np.diff(taps.astype(np.int64)) > 0) and int(taps[-1]) < 19937
    if n > (1 << 26):
        return  # too far to walk on the CPU; the polynomial arithmetic is the same code path
    bitgen = np.random.MT19937(5489)
    # std::mt19937(seed) and numpy's legacy seeding agree for init_genrand
    bitgen._legacy_seeding(5489)
    span = 19937 + 623
    words = bitgen.random_raw(n + span + 700).astype(np.uint32)
    for a in (1, 5, 333):
        acc = np.zeros(624, dtype=np.uint32)
        for i in taps.astype(np.int64):
            acc ^= words[a + i:a + i + 624]
        np.testing.assert_array_equal(acc, words[a + n:a + n + 624])


def test_pybind_binding_builds_and_has_no_cpu_fallback():
    """myfm_b200/csrc/pybind_binding.cpp compiles against the C ABI and, without a GPU, fails like the ctypes path."""
    from myfm_b200 import _lib
    from myfm_b200._myfm import ConfigBuilder
    from myfm_b200.csrc import build

    build.build_pybind()
    from myfm_b200 import _myfm_pybind

    assert _myfm_pybind.device_count() == _lib.device_count()
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    X, y = toy_matrix()
    cfg = ConfigBuilder().set_identical_groups(X.shape[1]).set_n_iter(2).set_n_kept_samples(2).build()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _myfm_pybind.create_train_fm(2, 0.1, X, [], y, 1, cfg, lambda *a: False)
