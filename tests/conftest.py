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
    """The product's CUDA library through its C-ABI binding; built if missing (nvcc, no GPU needed to build).
    LIINIT_GPU_TESTS_ON_EMUL=1 (developer pre-flight, never set by the suite): the bodies of `-m gpu` tests that only use the map / scan
    entry points run against the CPU build of the library (tests/emul) instead -- a dry run of a GPU test before GPU time is spent on it."""
    if os.environ.get("LIINIT_GPU_TESTS_ON_EMUL") == "1":
        import types
        import liinit_emul as le

        def make(ds, **kw):
            kw.pop("device_id", None)
            return le.EmulGpu(ds, **kw)
        return types.SimpleNamespace(LiInitGpu=make, LiInitError=le.EmulError, load=le.load)
    from lidar_imu_init_b200 import _build, capi
    _build.build_gpu()
    capi.load()
    return capi
