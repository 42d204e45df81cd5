"""Row-sharded data parallelism: one process per GPU (``torchrun``), every rank holds a contiguous
range of the training rows and a replica of the model; the engine sums the per-column sufficient
statistics of every dependency level with one NCCL all-reduce.

``torch.distributed`` is plumbing only (rendezvous, two small host-side collectives at setup); it
works with the ``gloo`` backend too, which is how the CPU test-suite covers this module.

    import torch.distributed as dist, myfm_b200
    from myfm_b200 import distributed as mdist
    dist.init_process_group("nccl")
    X_local, y_local, ctx = mdist.shard(X, y)            # or bring your own shard + mdist.context()
    with ctx.options(dtype="f32"):
        fm = myfm_b200.MyFMRegressor(rank=32).fit(X_local, y_local, group_shapes=...)
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass
from typing import Iterator, Optional, Tuple

import numpy as np
from scipy import sparse as sps

from . import _lib
from .options import engine_options


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of `rank`: the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_levels(X) -> np.ndarray:
    """Dependency level of every column of this shard (myfm_level_schedule)."""
    csr = _lib.CsrHolder(X)
    level = np.zeros(max(1, csr.shape[1]), dtype=np.int32)
    n_levels = C.c_int32()
    _lib.check(_lib.lib().myfm_level_schedule(C.byref(csr.struct), _lib.ptr(level, C.c_int32), C.byref(n_levels)))
    return level[:csr.shape[1]]


def relax_levels(X, level: np.ndarray) -> Tuple[np.ndarray, bool]:
    """One local relaxation with `level` as lower bounds (myfm_level_relax)."""
    csr = _lib.CsrHolder(X)
    out = np.ascontiguousarray(level, dtype=np.int32).copy()
    if out.shape[0] == 0:
        return out, False
    n_levels, changed = C.c_int32(), C.c_int32()
    _lib.check(_lib.lib().myfm_level_relax(C.byref(csr.struct), _lib.ptr(out, C.c_int32), C.byref(n_levels),
                                           C.byref(changed)))
    return out, bool(changed.value)


def agree_on_levels(X_local, group=None, max_rounds: int = 1000) -> np.ndarray:
    """The dependency-level schedule of the GLOBAL matrix, computed from row shards.

    level(j) = 1 + max level of an earlier column sharing a row with j; a shard only sees the
    conflicts of its own rows.  Every rank relaxes locally, the ranks take the element-wise
    maximum, and the two steps repeat until no rank changes anything: the result is the least
    fixed point, i.e. exactly what one process computes on the whole matrix.  One-hot fields
    converge in the first round."""
    import torch
    import torch.distributed as dist

    level = local_levels(X_local)
    for _ in range(max_rounds):
        t = torch.from_numpy(level.astype(np.int32))
        dev = _collective_device(group)
        t = t.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        level, changed = relax_levels(X_local, t.cpu().numpy())
        flag = torch.tensor([int(changed)], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            return level
    raise RuntimeError("dependency-level consensus did not converge")


def _collective_device(group=None):
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def _fresh_unique_id(group=None) -> bytes:
    """Collective: rank 0 draws an ncclUniqueId, every rank receives it.  One id serves ONE
    communicator, i.e. one trainer."""
    import torch.distributed as dist

    box = [None]
    if dist.get_rank(group) == 0:
        buf = (C.c_char * 128)()
        _lib.check(_lib.lib().myfm_nccl_unique_id(C.cast(buf, C.c_void_p)))
        box[0] = bytes(buf.raw)
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


@dataclass
class ShardContext:
    world_size: int
    rank: int
    row_offset: int
    n_rows_global: int
    nccl_unique_id: Optional[bytes]
    column_level: np.ndarray
    group: object = None
    rows: Optional[np.ndarray] = None  # global indices of this rank's rows (set by shard())

    @contextlib.contextmanager
    def options(self, **kwargs) -> Iterator[None]:
        """engine_options(...) carrying this shard's description.  Every trainer created under it uses the
        SAME NCCL communicator (the engine keeps one per ncclUniqueId and process: creating a communicator
        and running its first collective costs from 0.6 s on 2 GPUs to seconds on 8), so trainers of one
        context must be created, stepped and dropped in the same order on every rank.  `fresh_communicator()`
        gives the context a new id (collective) when an independent communicator is wanted."""
        with engine_options(world_size=self.world_size, rank=self.rank, row_offset=self.row_offset,
                            n_rows_global=self.n_rows_global, nccl_unique_id=self.nccl_unique_id,
                            column_level=self.column_level, row_ids=self.rows, **kwargs):
            yield

    def fresh_communicator(self) -> None:
        """Collective: the next trainers of this context get a new communicator."""
        if self.nccl_unique_id is not None:
            self.nccl_unique_id = _fresh_unique_id(self.group)


def context(X_local, row_offset: int, n_rows_global: int, group=None, with_nccl: bool = True) -> ShardContext:
    """Collective: agrees on the level schedule and shares rank 0's ncclUniqueId."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    levels = agree_on_levels(X_local, group)
    uid = _fresh_unique_id(group) if with_nccl and world > 1 else None
    return ShardContext(world, rank, int(row_offset), int(n_rows_global), uid, levels, group)


def partition_by_key(key: np.ndarray, n_keys: int, world_size: int) -> np.ndarray:
    """Rank of every row given the row's key in [0, n_keys) (its first-field column): keys are dealt
    out in serpentine order of their row counts (see `column_partition`)."""
    counts = np.bincount(key, minlength=n_keys)
    order = np.argsort(-counts, kind="stable")
    pos = np.arange(order.shape[0])
    lap, lane = pos // world_size, pos % world_size
    rank_sorted = np.where(lap % 2 == 0, lane, world_size - 1 - lane)
    rank_of_key = np.empty(order.shape[0], dtype=np.int32)
    rank_of_key[order] = rank_sorted
    return rank_of_key[key]


def column_partition(X: sps.csr_matrix, world_size: int) -> np.ndarray:
    """Rank of every row when rows are dealt out by their FIRST column (for one-hot / categorical
    tables: by the category of the first field, e.g. by user).  Every first-field column then has
    all its rows on one rank, which lets the engine sweep that field without any exchange between
    the GPUs.  Columns are dealt in serpentine order of their row counts, so
    every rank gets the same number of columns (the streaming pass costs per column as well as
    per row) and, to a fraction of a per cent, the same number of rows."""
    n = X.shape[0]
    key = np.full(n, -1, dtype=np.int64)
    has = np.diff(X.indptr) > 0
    key[has] = X.indices[X.indptr[:-1][has]]
    counts = np.bincount(key + 1, minlength=X.shape[1] + 1)  # slot 0: rows without entries
    order = np.argsort(-counts, kind="stable")               # heaviest column first
    pos = np.arange(order.shape[0])
    lap, lane = pos // world_size, pos % world_size
    rank_sorted = np.where(lap % 2 == 0, lane, world_size - 1 - lane)
    rank_of_key = np.empty(order.shape[0], dtype=np.int64)
    rank_of_key[order] = rank_sorted
    return rank_of_key[key + 1]


def shard(X, y, group=None, with_nccl: bool = True, partition: str = "column"):
    """Takes this rank's rows of a dataset every rank holds in full.

    partition="column" (default): rows are dealt out by their first column (`column_partition`);
    partition="rows": contiguous row ranges (`shard_bounds`).  Either way the engine sums the
    per-column statistics over the ranks wherever a column's rows are spread over several of them."""
    import torch.distributed as dist

    X = sps.csr_matrix(X)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if partition == "rows":
        lo, hi = shard_bounds(X.shape[0], world, rank)
        rows = np.arange(lo, hi)
    elif partition == "column":
        X.sort_indices()
        rows = np.flatnonzero(column_partition(X, world) == rank)
    else:
        raise ValueError("partition must be 'column' or 'rows'")
    X_local, y_local = X[rows], np.asarray(y)[rows]
    ctx = context(X_local, int(rows[0]) if rows.size else 0, X.shape[0], group, with_nccl)
    ctx.rows = rows
    return X_local, y_local, ctx
