"""BASELINE.json configurations at (or near) their stated sizes, through the C-ABI on the B200, against the oracle (the verbatim
ikd-Tree when oracle/_ref is present): C2 whole-frame accumulators, C3 (130k x 10M), C4 (grow / box-delete cycle: live set and
searches after every step), C5 (2M-point frame)."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu


def _bk(orc):
    return 1 if orc.has_ikd() else 0


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _threads():
    import os
    return os.cpu_count() or 1


def test_c2_full_size_accumulators_match_oracle(gpu_lib, oracle_mod):
    """The bench workload itself: HtH / Htr / m of the whole 240k-point frame vs the 5M-point map against the oracle (all host
    threads), not a sample; bench.py emits the same comparison as `parity` from its cpu_baseline leg."""
    c = scenes.make_config("C2")
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=6_000_000, max_scan_points=250_000)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(c["body_xyz"])
    Ho, bo, mo = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True, nthreads=_threads())
    assert m == mo
    assert _rel(H, Ho) <= 1e-9 and _rel(b, bo) <= 1e-9
    st, so = g.scan_state(), sc.get()
    assert np.array_equal(st["selected"], so["selected"]) and np.array_equal(st["near_cnt"], so["near_cnt"])
    assert np.array_equal(st["near_xyz"], so["near_xyz"])
    sel = so["selected"].astype(bool)
    assert np.array_equal(st["normvec"][sel], so["normvec"][sel])
    # the 12-column pass on the same frame (LIO mode after the hand-over)
    ci = scenes.make_config("C2", N=60000, imu_en=True)
    pi = ci["pose_init"]
    g.scan_upload(ci["body_xyz"])
    H12, b12, m12, _ = g.icp_iterate(pi.rot_end, pi.pos_end, pi.R_LI, pi.T_LI, True, True)
    sci = oracle_mod.OracleScan(ci["body_xyz"])
    Ho12, bo12, mo12 = sci.iterate(om, pi.rot_end, pi.pos_end, pi.R_LI, pi.T_LI, True, True, nthreads=_threads())
    assert m12 == mo12 and _rel(H12, Ho12) <= 1e-9 and _rel(b12, bo12) <= 1e-9
    g.close()


def test_c3_velodyne_shape(gpu_lib, oracle_mod):
    """C3: 130k-point spinning scan (det_range 100 m) against a 10M-point map; whole-frame accumulators and per-point state."""
    n0, m0, ds, det = scenes.CONFIGS["C3"]
    c = scenes.make_config("C3")
    assert len(c["body_xyz"]) == n0 and len(c["map_xyz"]) == m0
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(ds, max_map_points=12_000_000, max_scan_points=140_000)
    g.map_build(c["map_xyz"])
    assert g.map_validnum() == m0
    g.scan_upload(c["body_xyz"])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    om = oracle_mod.OracleMap(ds, _bk(oracle_mod))
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(c["body_xyz"])
    Ho, bo, mo = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True, nthreads=_threads())
    assert m == mo and m > 0.8 * n0
    assert _rel(H, Ho) <= 1e-9 and _rel(b, bo) <= 1e-9
    st, so = g.scan_state(), sc.get()
    assert np.array_equal(st["near_cnt"], so["near_cnt"]) and np.array_equal(st["near_xyz"], so["near_xyz"])
    assert np.array_equal(st["selected"], so["selected"])
    g.close()


def test_c4_growth_and_delete_cycle(gpu_lib, oracle_mod):
    """C4's path at a size the CPU tree can follow step by step: Build, then scans walking down a hall -- search pass, map_incremental
    (downsample + plain inserts) and a trailing Delete_Point_Boxes per scan. After every step: live count, live set and exact 5-NN of
    the next scan equal to the verbatim ikd-Tree's."""
    ds = 0.15
    scene = scenes.box_scene(120.0, 20.0, 6.0)
    mp = scenes.map_points(scene, ds, None, seed=5)
    mp = mp[mp[:, 0] < 30.0]                       # the map starts with the first 30 m of the hall
    g = gpu_lib.LiInitGpu(ds, max_map_points=2_000_000, max_scan_points=40_000)
    om = oracle_mod.OracleMap(ds, _bk(oracle_mod))
    g.map_build(mp)
    om.build(mp)
    R_LI, T_LI = scenes.identity_extrinsic()
    for k in range(6):
        gt = scenes.default_sensor_pose(scene, R_LI, T_LI)
        gt.pos_end[0] = 15.0 + 12.0 * k
        body = scenes.scan_points(scene, gt, 26_000, seed=40 + k, det_range=35.0, sigma=0.01, open_air_frac=0.01)
        init = scenes.perturb_pose(gt, 60 + k, dtheta_deg=0.2, dpos=0.03)
        g.scan_upload(body)
        sc = oracle_mod.OracleScan(body)
        H, b, m, _ = g.icp_iterate(init.rot_end, init.pos_end, R_LI, T_LI, False, True)
        Ho, bo, mo = sc.iterate(om, init.rot_end, init.pos_end, R_LI, T_LI, False, True)
        assert m == mo and _rel(H, Ho) <= 1e-9 and _rel(b, bo) <= 1e-9, k
        st, so = g.scan_state(), sc.get()
        assert np.array_equal(st["near_cnt"], so["near_cnt"]) and np.array_equal(st["near_xyz"], so["near_xyz"]), k
        na, nn = g.map_incremental(gt.rot_end, gt.pos_end, R_LI, T_LI, ds)
        _, oa, on, _ = sc.map_incremental(om, gt.rot_end, gt.pos_end, R_LI, T_LI, ds)
        assert (na, nn) == (oa, on) and g.map_validnum() == om.validnum(), k
        if k >= 2:   # the local map follows the sensor: drop what lies more than 30 m behind it (lasermap_fov_segment's boxes, wired)
            boxes = np.array([[-1.0, -1.0, -1.0, gt.pos_end[0] - 30.0, scene.W + 1.0, scene.H + 1.0]], np.float32)
            assert g.map_delete_boxes(boxes) == om.delete_boxes(boxes), k
            assert g.map_validnum() == om.validnum(), k
    assert set(map(bytes, g.map_download())) == set(map(bytes, om.flatten()))
    g.close()


def test_c5_two_million_point_frame(gpu_lib, oracle_mod):
    """C5 on one GPU: the 2M-point frame against the 5M-point map. Linearity over an uneven cut (what the rank slots rely on) and a
    sampled exact comparison with the oracle."""
    n0, m0, ds, det = scenes.CONFIGS["C5"]
    c = scenes.make_config("C5")
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(ds, max_map_points=6_000_000, max_scan_points=n0 + 16)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    st = g.scan_state()
    assert m == int(st["selected"].sum()) and m > 0.9 * n0
    acc, bcc, mm = np.zeros((12, 12)), np.zeros(12), 0
    cuts = [0, 700_001, 1_250_000, n0]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        g.scan_upload(c["body_xyz"][lo:hi])
        Hp, bp, mp_, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
        acc += Hp
        bcc += bp
        mm += mp_
    assert mm == m and _rel(acc, H) <= 1e-11 and _rel(bcc, b) <= 1e-9
    om = oracle_mod.OracleMap(ds, _bk(oracle_mod))
    om.build(c["map_xyz"])
    idx = np.sort(np.random.default_rng(3).permutation(n0)[:30000])
    sc = oracle_mod.OracleScan(c["body_xyz"][idx])
    sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True, nthreads=_threads())
    so = sc.get()
    assert np.array_equal(so["world"], st["world"][idx])
    assert np.array_equal(so["near_cnt"], st["near_cnt"][idx]) and np.array_equal(so["near_xyz"], st["near_xyz"][idx])
    assert np.array_equal(so["selected"], st["selected"][idx])
    sel = so["selected"].astype(bool)
    assert np.array_equal(so["normvec"][sel], st["normvec"][idx][sel])
    g.close()
