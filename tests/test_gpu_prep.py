"""Device-side data preparation (csrc/prep_device.cuh: row order, permuted CSR, CSC and field arrays built with
radix sorts on the GPU) against the host preparation (csrc/host_data.hpp): the same chain, bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sps

from helpers import fields_like, movielens_like
from test_gpu_parity import make_pair, run_chain_parity

pytestmark = pytest.mark.gpu


def _state(t):
    w0, w, V, _ = t.get_fm()
    return w0, w, V, t.get_e()


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("shape", ["two-unit", "three-weighted", "five-unit"])
def test_device_preparation_equals_host_preparation(engine, oracle, monkeypatch, dtype, shape):
    if shape == "two-unit":
        X, y, gs = movielens_like(30011, 500, 90, 3, seed=41)
    elif shape == "three-weighted":
        X, y, gs = fields_like(20000, [300, 40, 25], 3, seed=42, unit=False)
    else:
        X, y, gs = fields_like(12000, [120, 60, 30, 12, 7], 3, seed=43, unit=True)
    monkeypatch.delenv("MYFM_HOST_SETUP", raising=False)
    a, _ = make_pair(engine, oracle, X, y, 5, dtype, group_shapes=gs)
    monkeypatch.setenv("MYFM_HOST_SETUP", "1")
    b, _ = make_pair(engine, oracle, X, y, 5, dtype, group_shapes=gs)
    assert a.sweep_path() == b.sweep_path() == 1
    for it in range(4):
        a.step(1)
        b.step(1)
        for xa, xb in zip(_state(a), _state(b)):
            np.testing.assert_array_equal(xa, xb, err_msg=f"sweep {it}")


def test_device_preparation_against_oracle_and_fallbacks(engine, oracle, monkeypatch):
    """The device path against the oracle; inputs it declines (fields whose column ranges do not ascend with
    the position, ragged rows) are prepared on the host and still match the oracle."""
    monkeypatch.delenv("MYFM_HOST_SETUP", raising=False)
    X, y, gs = fields_like(9000, [200, 50, 20], 3, seed=44, unit=False)
    t, chain = make_pair(engine, oracle, X, y, 4, "f64", group_shapes=gs)
    run_chain_parity(t, chain, "f64", 4)
    # the same table with its column blocks in reverse order: every row still has three entries, but the first
    # entry of a row now lies in the highest block
    perm = np.concatenate([np.arange(250, 270), np.arange(200, 250), np.arange(0, 200)])
    Xr = sps.csr_matrix(X[:, perm])
    Xr.sort_indices()
    t, chain = make_pair(engine, oracle, Xr, y, 4, "f64", group_shapes=[20, 50, 200])
    run_chain_parity(t, chain, "f64", 3)
    Xg = X.tolil()
    Xg[5, X[5].indices[1]] = 0  # a ragged row
    Xg = Xg.tocsr()
    Xg.eliminate_zeros()
    t, chain = make_pair(engine, oracle, Xg, y, 4, "f64", group_shapes=gs)
    run_chain_parity(t, chain, "f64", 3)
