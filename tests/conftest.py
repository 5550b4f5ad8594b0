import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = ROOT / "tests" / "golden"
    return {p.stem: np.load(p) for p in d.glob("*.npz")}


@pytest.fixture(scope="session")
def gpu_ctx():
    """The process-wide libqtrack context; GPU tests fail (not skip) if it cannot be created."""
    from lqr_quadcopter_test_b200 import _lib

    return _lib.get_context()
