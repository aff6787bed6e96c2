import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure). Built on demand; prefers oracle/_ref (verbatim ikd-Tree)."""
    from oracle import oracle as orc
    orc.load()
    return orc


@pytest.fixture(scope="session")
def gpu_lib():
    """The product's CUDA library through its C-ABI binding; built if missing (nvcc, no GPU needed to build)."""
    from lidar_imu_init_b200 import _build, capi
    _build.build_gpu()
    capi.load()
    return capi
