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
