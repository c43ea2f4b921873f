import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def up():
    """The product package; GPU tests fail loudly if the CUDA library is missing."""
    import upsp_b200
    assert os.path.exists(upsp_b200.LIB_PATH), "libupsp_gpu.so not built (python upsp-processing_b200/build.py)"
    return upsp_b200


@pytest.fixture(scope="session")
def gpu(up):
    n = up.device_count()
    assert n > 0, "GPU test selected but no CUDA device is visible"
    return n
