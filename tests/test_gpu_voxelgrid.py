"""Row N2: scan voxel-grid downsample on the device vs the oracle's PCL VoxelGrid restatement."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu


def _lexsort(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


@pytest.mark.parametrize("leaf", [0.05, 0.2, 1.0])
def test_voxel_grid_matches_oracle(gpu_lib, oracle_mod, leaf):
    c = scenes.make_config("C2", N=60000, M=15000, open_air_frac=0.01, order="shuffle")
    raw = c["body_xyz"].copy()
    raw[100] = [np.nan, 1, 1]          # non-finite points are skipped (PCL: !isFinite -> continue)
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=100000, max_scan_points=80000)
    n = g.scan_upload_raw(raw, leaf)
    want = oracle_mod.voxel_grid(raw, leaf)
    got = g.scan_body()
    assert n == len(want) == len(got)
    assert np.array_equal(got, want)   # bit-equal centroids, in PCL's output order (ascending leaf index)
    # run to run deterministic, including the order
    n2 = g.scan_upload_raw(raw, leaf)
    assert n2 == n and np.array_equal(g.scan_body(), got)
    g.close()


def test_downsampled_scan_feeds_the_icp_pass(gpu_lib, oracle_mod):
    c = scenes.make_config("C2", N=40000, M=150000, open_air_frac=0.01)
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=300000, max_scan_points=50000)
    g.map_build(c["map_xyz"])
    n = g.scan_upload_raw(c["body_xyz"], 0.25)
    body = g.scan_body()
    H, b, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    om = oracle_mod.OracleMap(c["ds"], 1 if oracle_mod.has_ikd() else 0)
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(body)
    Ho, bo, mo = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    assert m == mo and n == len(body) < 40000
    assert np.abs(H - Ho).max() <= 1e-9 * np.abs(Ho).max() and np.abs(b - bo).max() <= 1e-9 * np.abs(bo).max()
    g.close()


def test_voxel_grid_errors(gpu_lib):
    g = gpu_lib.LiInitGpu(0.15, max_map_points=1000, max_scan_points=100)
    with pytest.raises(gpu_lib.LiInitError):           # more leaves than max_scan_points
        g.scan_upload_raw(np.random.default_rng(0).uniform(0, 50, (5000, 3)).astype(np.float32), 0.1)
    with pytest.raises(gpu_lib.LiInitError):           # leaf index overflow ("Leaf size is too small")
        g.scan_upload_raw(np.array([[0, 0, 0], [1e3, 1e3, 1e3]], np.float32), 0.001)
    g.close()
