import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (oracle/), the checker of every parity test."""
    from oracle import binding

    binding.build()
    return binding


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its Python host layer; skips when no GPU is visible."""
    from myfm_b200 import _lib

    if _lib.device_count() == 0:
        pytest.skip("no CUDA device")
    import myfm_b200

    return myfm_b200


# myfm_trainer_sweep_path values (include/myfm_b200.h)
PATH_GENERAL, PATH_FIELD, PATH_TILE = 0, 1, 5


@pytest.fixture(params=["tile", "field"])
def two_field_path(request, monkeypatch):
    """Two-field tables take the field path (csrc/field_sweep.cuh) by default; MYFM_TILE_PATH=1 selects the
    tile path (csrc/tile_sweep.cuh: row tiles staged in shared memory by TMA bulk copies) on one GPU.
    Yields the sweep path the trainer must report."""
    if request.param == "tile":
        monkeypatch.setenv("MYFM_TILE_PATH", "1")
        return PATH_TILE
    monkeypatch.delenv("MYFM_TILE_PATH", raising=False)
    return PATH_FIELD
