"""The node-side patch of INTEGRATION.md as a real C++ translation unit (tests/cpp/dropin_node.cpp): liinit_adapter.hpp +
liinit_host.h instantiated with the reference's 48-byte pcl::PointXYZINormal and Eigen::aligned_allocator. Compiled on the CPU;
on a GPU its per-scan sequence is compared with the ctypes path on the same inputs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_node")


def _build():
    from lidar_imu_init_b200 import _build
    _build.build_gpu()
    _build.build_host()
    src = os.path.join(ROOT, "tests", "cpp", "dropin_node.cpp")
    lib = os.path.join(ROOT, "lidar_imu_init_b200")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(_build.GPU_LIB), os.path.getmtime(_build.HOST_LIB)):
        return EXE
    gxx = shutil.which("g++") or "/usr/bin/g++"
    subprocess.check_call([gxx, "-O1", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(lib, "csrc", "host"),
                           "-I", os.path.join(ROOT, "oracle", "shim"), src, "-o", EXE, "-L", lib, "-lliinit_host", "-lliinit_gpu",
                           "-Wl,-rpath," + lib])
    return EXE


def test_dropin_translation_unit_compiles():
    """-std=c++14 like the node (CMakeLists.txt:8), -Wall -Werror: the adapter and the host header are usable as shipped."""
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_dropin_node_sequence_equals_ctypes_path(tmp_path, gpu_lib):
    from lidar_imu_init_b200 import host, scenes
    exe = _build()
    c = scenes.make_config("C2", N=20000, M=150000, open_air_frac=0.02)
    p = c["pose_init"]
    st0 = host.state_from_pose(p.rot_end, p.pos_end, p.R_LI, p.T_LI)
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(np.array([len(c["map_xyz"]), len(c["body_xyz"]), 0], np.int32).tobytes())
        f.write(np.ascontiguousarray(c["map_xyz"], np.float32).tobytes())
        f.write(np.ascontiguousarray(c["body_xyz"], np.float32).tobytes())
        f.write(np.ascontiguousarray(st0, np.float64).tobytes())
        f.write(np.array([c["ds"]], np.float64).tobytes())
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(inp), str(out)])
    raw = open(out, "rb").read()
    st_cpp = np.frombuffer(raw[:612 * 8], np.float64)
    tail = np.frombuffer(raw[612 * 8:612 * 8 + 32], np.int32)
    nx = np.frombuffer(raw[612 * 8 + 32:612 * 8 + 32 + 60], np.float32).reshape(5, 3)
    nd = np.frombuffer(raw[612 * 8 + 92:612 * 8 + 112], np.float32)
    # the same sequence through ctypes
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=len(c["map_xyz"]) * 2 + 100000, max_scan_points=len(c["body_xyz"]) + 16)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    st_py, info = host.scan_update(g, st0, 5, False)
    Rr, pp, RL, TL = host.state_pose(st_py)
    na, nn = g.map_incremental(Rr, pp, RL, TL, c["ds"])
    q = c["map_xyz"][len(c["map_xyz"]) // 2].copy()
    q[0] += np.float32(0.05)
    gx, gd, gc = g.nearest_search(q[None, :])
    assert np.array_equal(st_cpp, st_py)                       # bit-identical posterior state and covariance
    assert list(tail) == [info["iterations"], info["search_passes"], info["effect_feat_num"], na, nn, g.map_size(), g.map_validnum(), int(gc[0])]
    assert np.array_equal(nx[:gc[0]], gx[0][:gc[0]]) and np.array_equal(nd[:gc[0]], gd[0][:gc[0]])
    g.close()


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_dropin_node_two_ranks_cpp_only(tmp_path, gpu_lib):
    """The multi-GPU block of INTEGRATION.md section 5 with NO Python in the ranks: tests/cpp/dropin_node forks one process per GPU, rank 0
    draws the communicator id through the C-ABI, the others read it from a file, every rank runs liinit_scan_update (C++ IESKF loop; the
    sum over the ranks happens inside liinit_icp_iterate) and liinit_map_incremental. Both ranks must end bit-identical, and equal to the
    single-process run up to the association of the two-rank sum."""
    from lidar_imu_init_b200 import host, scenes
    exe = _build()
    c = scenes.make_config("C2", N=20000, M=150000, open_air_frac=0.02)
    p = c["pose_init"]
    st0 = host.state_from_pose(p.rot_end, p.pos_end, p.R_LI, p.T_LI)
    inp = tmp_path / "in.bin"
    with open(inp, "wb") as f:
        f.write(np.array([len(c["map_xyz"]), len(c["body_xyz"]), 0], np.int32).tobytes())
        f.write(np.ascontiguousarray(c["map_xyz"], np.float32).tobytes())
        f.write(np.ascontiguousarray(c["body_xyz"], np.float32).tobytes())
        f.write(np.ascontiguousarray(st0, np.float64).tobytes())
        f.write(np.array([c["ds"]], np.float64).tobytes())
    subprocess.check_call([exe, str(inp), str(tmp_path / "one.bin")])
    subprocess.check_call([exe, str(inp), str(tmp_path / "two.bin"), "2"], timeout=300)
    one = open(tmp_path / "one.bin", "rb").read()
    r0, r1 = open(str(tmp_path / "two.bin") + ".0", "rb").read(), open(str(tmp_path / "two.bin") + ".1", "rb").read()
    assert r0 == r1                                             # every rank: same state, covariance, counters, neighbours
    s1, s2 = np.frombuffer(one[:612 * 8], np.float64), np.frombuffer(r0[:612 * 8], np.float64)
    assert np.abs(s1[:24] - s2[:24]).max() <= 1e-6             # (two-rank sum associates differently: ~1e-8 after five 24 x 24 solves)
    t1, t2 = np.frombuffer(one[612 * 8:612 * 8 + 32], np.int32), np.frombuffer(r0[612 * 8:612 * 8 + 32], np.int32)
    assert t1[0] == t2[0] and t1[1] == t2[1] and abs(int(t1[2]) - int(t2[2])) <= 2 and np.abs(t1[3:7] - t2[3:7]).max() <= 3
