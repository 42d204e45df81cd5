"""The tile path of the column sweeps (myfm_b200/csrc/tile_sweep.cuh: row tiles staged in shared memory
by TMA bulk copies; pending update, first-field sweep and last-field statistics in one pass, partial sums
folded in tile order) against the oracle, against the field path, and for run-to-run reproducibility."""
import numpy as np
import pytest

from helpers import fields_like, movielens_like
from test_gpu_parity import make_pair, run_chain_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("tile_rows", [None, 2001])
@pytest.mark.parametrize("unit", [True, False])
def test_tile_path_against_oracle(engine, oracle, monkeypatch, dtype, tile_rows, unit):
    """Two fields, unit and weighted values; default tiles and 2000-row tiles (dozens of tiles, tiles that
    start and end on odd rows: the bulk store covers the 16-byte aligned interior, single rows go with
    ordinary stores)."""
    monkeypatch.setenv("MYFM_TILE_PATH", "1")
    if tile_rows:
        monkeypatch.setenv("MYFM_TILE_ROWS", str(tile_rows))
    X, y, gs = fields_like(9001, [333, 47], 3, seed=21, unit=unit, zipf=0.9)
    t, chain = make_pair(engine, oracle, X, y, 5, dtype, group_shapes=gs)
    assert t.sweep_path() == 5
    run_chain_parity(t, chain, dtype, 6 if dtype == "f64" else 3)


def test_tile_path_equals_field_path(engine, oracle, monkeypatch):
    """Same chain through the tile path and the field path: summation order is the only difference."""
    X, y, gs = movielens_like(50001, 900, 120, 4, seed=22, zipf=(0.5, 0.9))
    monkeypatch.setenv("MYFM_TILE_ROWS", "2000")
    monkeypatch.setenv("MYFM_TILE_PATH", "1")
    a, _ = make_pair(engine, oracle, X, y, 8, "f64", group_shapes=gs)
    monkeypatch.delenv("MYFM_TILE_PATH")
    b, _ = make_pair(engine, oracle, X, y, 8, "f64", group_shapes=gs)
    assert (a.sweep_path(), b.sweep_path()) == (5, 1)
    for it in range(5):
        a.step(1)
        b.step(1)
        for xa, xb in zip(a.get_fm()[:3], b.get_fm()[:3]):
            np.testing.assert_allclose(xa, xb, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(a.get_e(), b.get_e(), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(a.get_q(), b.get_q(), rtol=1e-9, atol=1e-9)


def test_tile_path_is_bit_reproducible(engine, oracle, monkeypatch):
    """Work is handed out dynamically inside a tile, but every sum has a fixed order: two runs agree bit for bit."""
    X, y, gs = movielens_like(40000, 500, 200, 4, seed=23)
    monkeypatch.setenv("MYFM_TILE_ROWS", "1500")
    monkeypatch.setenv("MYFM_TILE_PATH", "1")
    runs = []
    for _ in range(2):
        t, _ = make_pair(engine, oracle, X, y, 6, "f32", group_shapes=gs)
        t.step(4)
        runs.append((t.get_fm(), t.get_e()))
    (fa, ea), (fb, eb) = runs
    assert fa[0] == fb[0]
    np.testing.assert_array_equal(fa[1], fb[1])
    np.testing.assert_array_equal(fa[2], fb[2])
    np.testing.assert_array_equal(ea, eb)


def test_empty_categories_two_fields(engine, oracle, monkeypatch):
    """Categories without rows in the first and in the last field are drawn from the prior (once)."""
    monkeypatch.setenv("MYFM_TILE_PATH", "1")
    X, y, gs = fields_like(4000, [40, 16], 2, seed=24)
    X = X.tolil()
    for j, repl in ((3, 0), (45, 41)):
        for r in X[:, j].nonzero()[0]:
            X[r, j] = 0
            X[r, repl] = 1.0
    X = X.tocsr()
    X.eliminate_zeros()
    assert X.getnnz(axis=0)[3] == 0 and X.getnnz(axis=0)[45] == 0
    t, chain = make_pair(engine, oracle, X, y, 3, "f64", group_shapes=gs)
    assert t.sweep_path() == 5
    run_chain_parity(t, chain, "f64", 4)
