"""CPU tests of the product boundary: the C-ABI library builds, loads and exports every symbol that
include/liinit_gpu.h declares; without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "liinit_gpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(liinit_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported(gpu_lib):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(gpu_lib.lib_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/liinit_gpu.h but not exported"
    assert set(gpu_lib.SYMBOLS) == set(names)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under lidar_imu_init_b200/ may import, link or load it."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "lidar_imu_init_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/_ref|#\s*include\s*[<\"][^>\"]*(oracle|ikd_Tree)", txt, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_product_does_not_reference_the_cpu_checkers():
    """tests/emul (the kernels run on the CPU through a SIMT shim) is test infrastructure as well: no Python module of the package may
    load it or its libraries, and the built product library must not contain the shim (the sources only carry LI_SIMT_EMUL guards)."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "lidar_imu_init_b200")):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"emul|simt_shim|cuda_shim", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    build = open(os.path.join(ROOT, "lidar_imu_init_b200", "_build.py")).read()
    assert "LI_SIMT_EMUL" not in build and "tests" not in build


def test_no_gpu_fails_loudly(gpu_lib):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("a GPU is present")
    with pytest.raises(gpu_lib.LiInitError) as e:
        gpu_lib.LiInitGpu()
    assert e.value.code == -2   # LIINIT_ERR_CUDA: no device, no fallback


def test_config_struct_matches_header(gpu_lib):
    assert ctypes.sizeof(gpu_lib.Config) == 4 * (7 + 8)
