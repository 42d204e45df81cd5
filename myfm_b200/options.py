"""Engine options that have no counterpart in the reference API.

The estimator signatures stay exactly the reference's, so the knobs of the CUDA engine live
here: ``myfm_b200.engine_options(dtype="f32")`` as a context manager or a plain setter, with
``MYFM_B200_DTYPE`` / ``MYFM_B200_RNG`` / ``MYFM_B200_DEVICE`` (or ``MYFM_DTYPE`` / ``MYFM_RNG`` /
``MYFM_DEVICE``) as environment defaults.

dtype  "f64" (default; what the reference ships, cpp_source/bind.cpp) or "f32" (the reference's
       bind_float.cpp instantiation; the fast path the benchmarks use).
rng    "mt19937": the libstdc++ stream of the reference, draw for draw (same seed, comparable
       chain); "philox": device counter-based RNG (statistically equivalent chain).
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass, replace
from typing import Iterator, Optional


def _env(name: str, default: str) -> str:
    """MYFM_B200_<name>, or the shorter MYFM_<name>."""
    return os.environ.get("MYFM_B200_" + name, os.environ.get("MYFM_" + name, default))


@dataclass(frozen=True)
class EngineOptions:
    dtype: str = _env("DTYPE", "f64")
    rng: str = _env("RNG", "mt19937")
    device: int = int(_env("DEVICE", os.environ.get("LOCAL_RANK", "0")))
    # row-sharded data parallelism (set by myfm_b200.distributed)
    world_size: int = 1
    rank: int = 0
    row_offset: int = 0
    n_rows_global: int = 0
    nccl_unique_id: Optional[bytes] = None
    column_level: Optional[object] = None  # int32 array: dependency level of every main-table column
    row_ids: Optional[object] = None  # int64 array: global index of every row of this shard


_current = EngineOptions()


def get_options() -> EngineOptions:
    return _current


def set_options(**kwargs) -> EngineOptions:
    global _current
    opts = replace(_current, **kwargs)
    if opts.dtype not in ("f32", "f64"):
        raise ValueError("dtype must be 'f32' or 'f64'")
    if opts.rng not in ("mt19937", "philox"):
        raise ValueError("rng must be 'mt19937' or 'philox'")
    _current = opts
    return _current


@contextlib.contextmanager
def engine_options(**kwargs) -> Iterator[EngineOptions]:
    global _current
    saved = _current
    try:
        yield set_options(**kwargs)
    finally:
        _current = saved
