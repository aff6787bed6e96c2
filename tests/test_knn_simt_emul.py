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


REL = 1e-9


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("imu_en", [False, True])
def test_full_pass_on_the_cpu_matches_oracle(oracle_mod, imu_en):
    """k_knn_scan + k_icp_plane<imu, SEARCH> + k_icp_plane<imu, reuse> from the kernels' own source: fp64 Householder plane fit,
    gating, Jacobian rows, warp reduce-scatter, per-block partials, ticketed last-block reduction -- same bars as the GPU
    parity test (tests/test_gpu_parity.py::test_search_and_reuse_pass)."""
    c = scenes.make_config("C2", N=3000, M=60000, open_air_frac=0.02, imu_en=imu_en)
    p = c["pose_init"]
    p2 = scenes.perturb_pose(p, 77, dtheta_deg=0.05, dpos=0.01)
    m = ks.SimtMap(c["map_xyz"], c["ds"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    r = m.icp_pass(c["body_xyz"], p, imu_en, pose2=p2)
    Ho, bo, mo = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    so = osc.get()
    assert r["m"] == mo and mo > 2500
    assert np.array_equal(r["world"], so["world"]) and np.array_equal(r["near_cnt"], so["near_cnt"]) and np.array_equal(r["near_xyz"], so["near_xyz"])
    assert _relerr(r["H"], Ho) <= REL and _relerr(r["b"], bo) <= REL
    if not imu_en:
        assert np.all(r["H"][6:, :] == 0) and np.all(r["H"][:, 6:] == 0) and np.all(r["b"][6:] == 0)
    _, meas, _ = osc.get_H()
    assert abs(r["res_sq"] - float((meas ** 2).sum())) <= 1e-9 * max(r["res_sq"], 1e-30)
    # reuse pass at a moved pose (laserMapping.cpp:989-994): stored neighbours and flags
    Ho2, bo2, mo2 = osc.iterate(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    so2 = osc.get()
    assert r["m2"] == mo2 and _relerr(r["H2"], Ho2) <= REL and _relerr(r["b2"], bo2) <= REL
    assert np.array_equal(r["selected"], so2["selected"])
    sel = so2["selected"].astype(bool)
    assert np.array_equal(r["normvec"][sel], so2["normvec"][sel])     # f32 normal + residual: bit-equal
    m.close()


def test_full_pass_on_the_cpu_analytic_planar_scene(oracle_mod):
    """BASELINE config 1 (noise-free planes): the pass against the closed form -- residual = signed distance to the plane."""
    c = scenes.make_config("C1")
    p = c["pose_init"]
    m = ks.SimtMap(c["map_xyz"], c["ds"])
    r = m.icp_pass(c["body_xyz"], p, False)
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    Ho, bo, mo = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    # grid-aligned map: exact distance ties decide WHICH equidistant points are the 5 neighbours, so compare what does not
    # depend on that choice: the count, and that the planes found are mostly the scene's axis-aligned ones
    sel = r["selected"].astype(bool)
    n = r["normvec"][sel, :3]
    assert sel.sum() > 0.4 * len(sel) and abs(r["m"] - mo) <= 0.02 * mo, (int(sel.sum()), r["m"], mo)
    assert (np.abs(np.abs(n).max(axis=1) - 1.0) < 1e-4).mean() > 0.5   # away from the edges where two planes share the neighbours
    m.close()
