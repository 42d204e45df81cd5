"""Host-side logic of row-sharded training, covered on CPU with world_size-2 gloo process groups:
shard bounds, and the dependency-level consensus (the fixed point of local relaxation + MAX
all-reduce must equal the schedule one process computes on the whole matrix)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, case: str, out_dir: str, partition: str) -> None:
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from myfm_b200 import distributed as mdist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = _case(case)
        X_local, _, ctx = mdist.shard(X, np.zeros(X.shape[0]), with_nccl=False, partition=partition)
        assert ctx.world_size == world and ctx.rank == rank and ctx.n_rows_global == X.shape[0]
        np.save(os.path.join(out_dir, f"levels_{rank}.npy"), ctx.column_level)
        np.save(os.path.join(out_dir, f"rows_{rank}.npy"), ctx.rows)
        # the engine options a trainer of this shard is created under: the same description (and the same
        # communicator id: trainers of one context share it) every time the context is used
        from myfm_b200.options import get_options

        seen = []
        for _ in range(2):
            with ctx.options(dtype="f32"):
                o = get_options()
                seen.append((o.world_size, o.rank, o.n_rows_global, o.nccl_unique_id, o.dtype))
                assert np.array_equal(o.row_ids, ctx.rows) and np.array_equal(o.column_level, ctx.column_level)
        assert seen[0] == seen[1] == (world, rank, X.shape[0], None, "f32")
        ctx.fresh_communicator()  # no NCCL id in this context: a no-op
        assert ctx.nccl_unique_id is None
    finally:
        dist.destroy_process_group()


def _case(case: str) -> sps.csr_matrix:
    if case == "onehot":
        from helpers import movielens_like

        return movielens_like(3001, 80, 30, 2, seed=3)[0]
    rng = np.random.default_rng(7)
    if case == "chain":
        # a conflict chain that crosses the shard boundary: column j conflicts with j+1 through a
        # row that lives in the OTHER shard for every second link
        n = 12
        rows = []
        for j in range(n - 1):
            rows.append((j, j + 1))
        order = [i for i in range(0, n - 1, 2)] + [i for i in range(1, n - 1, 2)]
        data = sps.lil_matrix((n - 1, n))
        for r, link in enumerate(order):
            a, b = rows[link]
            data[r, a] = 1.0
            data[r, b] = 2.0
        return data.tocsr()
    return sps.random(400, 60, density=0.05, random_state=np.random.RandomState(5), format="csr")


@pytest.mark.parametrize("partition", ["rows", "column"])
@pytest.mark.parametrize("case", ["onehot", "chain", "random"])
def test_level_consensus_two_ranks(case, partition, tmp_path):
    import torch.multiprocessing as mp

    sys.path.insert(0, ROOT)
    from myfm_b200 import distributed as mdist

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path), partition), nprocs=world, join=True)
    X = _case(case)
    want = mdist.local_levels(X)  # one process, whole matrix
    got = [np.load(tmp_path / f"levels_{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(got[0], got[1])
    np.testing.assert_array_equal(got[0], want)
    rows = [np.load(tmp_path / f"rows_{r}.npy") for r in range(world)]
    assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(X.shape[0]))  # a partition of the rows
    if partition == "rows":
        assert rows[0][0] == 0 and rows[0][-1] + 1 == rows[1][0] and rows[1][-1] + 1 == X.shape[0]
    else:  # every first column has all its rows on one rank
        Xs = sps.csr_matrix(X)
        Xs.sort_indices()
        first = [set(Xs.indices[Xs.indptr[r]] for r in rr if Xs.indptr[r + 1] > Xs.indptr[r]) for rr in rows]
        assert not (first[0] & first[1])


def test_shard_bounds_cover_all_rows():
    sys.path.insert(0, ROOT)
    from myfm_b200.distributed import shard_bounds

    for n in (0, 1, 7, 10_000_054):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
