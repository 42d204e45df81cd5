"""Edge shapes through the C ABI against the oracle: a single row, one column per field, a first
field with one giant column (whole-CTA two-pass class), empty categories, wide ranks on both sides
of the tile forward pass's limit."""
import numpy as np
import pytest
import scipy.sparse as sps

from helpers import fields_like, movielens_like
from test_gpu_parity import make_pair, run_chain_parity

pytestmark = pytest.mark.gpu


def test_single_row(engine, oracle):
    X = sps.csr_matrix(np.asarray([[1.0, 0.0, 1.0]]))
    t, chain = make_pair(engine, oracle, X, np.asarray([0.7]), 2, "f64")
    run_chain_parity(t, chain, "f64", 3)


def test_one_column_fields_and_giant_first_column(engine, oracle):
    """Field 0 has ONE category holding all 20 000 rows (f64: above 4096 rows -> two-pass CTA class),
    the last field one category too (a one-entry shared-memory table)."""
    rng = np.random.default_rng(0)
    n = 20_000
    cols = np.stack([np.zeros(n, dtype=np.int64), 1 + rng.integers(0, 7, n), np.full(n, 8)], axis=1)
    X = sps.csr_matrix((rng.uniform(0.5, 1.5, 3 * n), cols.ravel(), np.arange(0, 3 * n + 1, 3)), shape=(n, 9))
    y = rng.normal(size=n)
    for dtype in ("f64", "f32"):
        t, chain = make_pair(engine, oracle, X, y, 3, dtype, group_shapes=[1, 7, 1])
        assert t.sweep_path() == 1
        run_chain_parity(t, chain, dtype, 3 if dtype == "f64" else 2)


def test_empty_categories_inside_fields(engine, oracle):
    """Categories without any row (drawn from the prior) in the first, a middle and the last field."""
    X, y, gs = fields_like(3000, [30, 12, 20], 2, seed=9)
    X = X.tolil()
    keep = np.ones(X.shape[1], dtype=bool)
    X = X.tocsc()
    for j in (3, 35, 61):  # one category of each field loses its rows
        rows = X[:, j].nonzero()[0]
        lo = 0 if j < 30 else (30 if j < 42 else 42)
        Xl = X.tolil()
        for r in rows:
            Xl[r, j] = 0
            Xl[r, lo] = 1.0
        X = Xl.tocsc()
    X = X.tocsr()
    X.eliminate_zeros()
    X.sum_duplicates()
    assert X.getnnz(axis=0)[3] == 0 and X.getnnz(axis=0)[61] == 0
    t, chain = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=gs)
    assert t.sweep_path() == 1
    run_chain_parity(t, chain, "f64", 4)


@pytest.mark.parametrize("rank", [64, 70])
def test_wide_ranks(engine, oracle, rank):
    X, y, gs = movielens_like(2000, 40, 20, 2, seed=5)
    t, chain = make_pair(engine, oracle, X, y, rank, "f64", group_shapes=gs)
    run_chain_parity(t, chain, "f64", 2)
