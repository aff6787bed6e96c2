"""The DEFAULT search kernel of the product (lockstep lane groups over whole bricks, knn_kernels.cuh: k_knn_scan ->
knn5_lockstep -> group scans / merges) executed on the CPU, unchanged, through tests/emul/simt_shim.h (one warp = 32 fibers,
every warp intrinsic a rendezvous) and compared bit for bit with the oracle (verbatim ikd-Tree when oracle/_ref is present).
The GPU parity tests stay the proof on the real hardware; this keeps the kernel's LOGIC under test where there is no GPU."""
import numpy as np
import pytest

import knn_simt_emul as ks
from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.skipif(not ks.available(), reason="g++ or the CUDA vector-type headers are missing")


def _world(body, p):
    return (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)


def _bk(orc):
    return 1 if orc.has_ikd() else 0


@pytest.fixture(scope="module")
def case():
    return scenes.make_config("C2", N=3000, M=60000, open_air_frac=0.02)


@pytest.mark.parametrize("G", [4, 8, 32])
@pytest.mark.parametrize("pose", ["pose_init", "pose_gt"])
def test_lockstep_kernel_matches_oracle(oracle_mod, case, G, pose):
    c, p = case, case[pose]
    m = ks.SimtMap(c["map_xyz"], c["ds"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    w, near, cnt = m.knn_scan(c["body_xyz"], p, G=G)
    q = _world(c["body_xyz"], p)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(w, q)                      # pointBodyToWorld: f64 math, f32 store
    assert np.array_equal(cnt, oc) and np.array_equal(near, ox)
    assert (cnt == 0).sum() >= 40 and (cnt == 5).sum() > 2500
    m.close()


@pytest.mark.parametrize("rho", [0.15, 1.0])
def test_lockstep_kernel_seed_radius_does_not_matter(oracle_mod, case, rho):
    c, p = case, case["pose_init"]
    m = ks.SimtMap(c["map_xyz"], c["ds"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    w, near, cnt = m.knn_scan(c["body_xyz"][:1500], p, rho=rho)
    ox, od, oc, _ = om.knn(_world(c["body_xyz"][:1500], p))
    assert np.array_equal(cnt, oc) and np.array_equal(near, ox)
    m.close()


def test_lockstep_kernel_large_coordinates_and_ragged_tail(oracle_mod):
    """6 km from the origin (coarse float cells) and a scan length that leaves idle groups in the last warp batch."""
    c = scenes.make_config("C2", N=1003, M=30000, open_air_frac=0.02)
    off = np.array([6000.0, -4500.0, 300.0])
    mp = (c["map_xyz"].astype(np.float64) + off).astype(np.float32)
    p = c["pose_init"].copy()
    p.pos_end = p.pos_end + off
    m = ks.SimtMap(mp, c["ds"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(mp)
    w, near, cnt = m.knn_scan(c["body_xyz"], p)
    ox, od, oc, _ = om.knn(w)
    assert np.array_equal(cnt, oc) and np.array_equal(near, ox)
    m.close()


def test_lockstep_kernel_non_finite_points_and_tiny_map(oracle_mod, case):
    c, p = case, case["pose_init"]
    body = c["body_xyz"][:300].copy()
    body[3] = [np.nan, 0, 0]
    body[77] = [np.inf, 1, 1]
    body[299] = [1e30, 0, 0]
    m = ks.SimtMap(c["map_xyz"], c["ds"])
    w, near, cnt = m.knn_scan(body, p)
    assert cnt[3] == 0 and cnt[77] == 0 and cnt[299] == 0 and (cnt == 5).sum() > 250
    m.close()
    tiny = ks.SimtMap(c["map_xyz"][:3], c["ds"])
    w, near, cnt = tiny.knn_scan(c["body_xyz"][:64], p)
    ox, od, oc, _ = oracle_mod.knn_bruteforce(c["map_xyz"][:3], w)
    assert np.array_equal(cnt, oc) and np.array_equal(near, ox)
    tiny.close()
