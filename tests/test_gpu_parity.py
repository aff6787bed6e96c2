"""GPU parity tests: the CUDA path through the C-ABI vs the CPU oracle on identical seeded inputs.

Bars: identical neighbour sets (up to exact distance ties), equal selected-point
count and flags, bit-equal f32 normal/residual, HtH / Htr relative error <= 1e-9 per pass.
"""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu

REL = 1e-9


def _world(body, p):
    return (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)


def _best_backend(orc):
    return 1 if orc.has_ikd() else 0


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def small_case():
    return scenes.make_config("C2", N=20000, M=200000, open_air_frac=0.02)


@pytest.mark.parametrize("group", [0, 8])
def test_knn_matches_oracle(gpu_lib, oracle_mod, small_case, group):
    c = small_case
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_group_lanes=group)
    g.map_build(c["map_xyz"])
    assert g.map_validnum() == len(c["map_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _best_backend(oracle_mod))
    om.build(c["map_xyz"])
    q = _world(c["body_xyz"], c["pose_init"])
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc)
    assert np.array_equal(gd, od), f"d2 mismatch at {np.argwhere(gd != od)[:5]}"
    assert np.array_equal(gx, ox)
    assert (gc == 0).sum() >= 300 and (gc == 5).sum() > 15000
    g.close()


@pytest.mark.parametrize("imu_en", [False, True])
@pytest.mark.parametrize("tile", [0, 2, 4, 8, 16, 32])
def test_search_and_reuse_pass(gpu_lib, oracle_mod, imu_en, tile):
    c = scenes.make_config("C2", N=20000, M=200000, open_air_frac=0.02, imu_en=imu_en)
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_group_lanes=tile)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _best_backend(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    # search pass
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    Ho, bo, mo = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    assert m == mo and m > 15000
    st = g.scan_state()
    so = osc.get()
    assert np.array_equal(st["world"], so["world"])
    assert np.array_equal(st["near_cnt"], so["near_cnt"])
    assert np.array_equal(st["near_xyz"], so["near_xyz"])
    assert np.array_equal(st["selected"], so["selected"])
    sel = so["selected"].astype(bool)
    assert np.array_equal(st["normvec"][sel], so["normvec"][sel])
    assert _relerr(H, Ho) <= REL and _relerr(b, bo) <= REL
    if not imu_en:
        assert np.all(H[6:, :] == 0) and np.all(H[:, 6:] == 0) and np.all(b[6:] == 0)
    _, meas, _ = osc.get_H()
    assert abs(rs - float((meas ** 2).sum())) <= 1e-9 * max(rs, 1e-30)
    # reuse pass at a moved pose (laserMapping.cpp:989-994)
    p2 = scenes.perturb_pose(p, 77, dtheta_deg=0.05, dpos=0.01)
    H2, b2, m2, _ = g.icp_iterate(p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    Ho2, bo2, mo2 = osc.iterate(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    assert m2 == mo2
    assert _relerr(H2, Ho2) <= REL and _relerr(b2, bo2) <= REL
    st2, so2 = g.scan_state(), osc.get()
    assert np.array_equal(st2["selected"], so2["selected"])
    # effect cloud (laserCloudOri / corr_normvect)
    ori, nv = g.scan_effect()
    sel2 = so2["selected"].astype(bool)
    assert np.array_equal(ori, c["body_xyz"][sel2]) and np.array_equal(nv, so2["normvec"][sel2])
    # run-to-run determinism of the reduction
    H3, b3, m3, _ = g.icp_iterate(p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    g.close()


def _setdiff_count(a, b):
    sa = set(map(bytes, np.ascontiguousarray(a, np.float32)))
    sb = set(map(bytes, np.ascontiguousarray(b, np.float32)))
    return len(sa - sb), len(sb - sa)


def test_add_points_matches_oracle(gpu_lib, oracle_mod):
    c = scenes.make_config("C2", N=30000, M=100000, open_air_frac=0.0)
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=600000, max_scan_points=50000)
    g.map_build(c["map_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _best_backend(oracle_mod))
    om.build(c["map_xyz"])
    new = _world(c["body_xyz"], c["pose_gt"])
    n1 = g.map_add_points(new[:20000], True)
    om.add_points(new[:20000], True)
    g.map_add_points(new[20000:], False)
    om.add_points(new[20000:], False)
    # second downsample batch hitting voxels that now hold several points
    g.map_add_points(new[15000:25000] + np.float32(0.01), True)
    om.add_points(new[15000:25000] + np.float32(0.01), True)
    gm, omap = g.map_download(), om.flatten()
    assert g.map_validnum() == om.validnum() == len(gm)
    only_g, only_o = _setdiff_count(gm, omap)
    assert only_g == 0 and only_o == 0, (only_g, only_o)
    assert n1 > 0
    g.close()


def test_map_incremental_matches_oracle(gpu_lib, oracle_mod):
    c = scenes.make_config("C2", N=20000, M=150000, open_air_frac=0.02)
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=600000, max_scan_points=50000)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _best_backend(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    gt = c["pose_gt"]
    na, nn = g.map_incremental(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"])
    _, oa, on, _ = osc.map_incremental(om, gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"])
    assert (na, nn) == (oa, on)
    gm, omap = g.map_download(), om.flatten()
    only_g, only_o = _setdiff_count(gm, omap)
    assert only_g == 0 and only_o == 0, (only_g, only_o)
    # and a second scan searches the updated map identically
    q = _world(c["body_xyz"][:5000], gt)
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.close()


@pytest.mark.parametrize("group", [0, 4, 8])
def test_seeded_later_search_pass_is_identical(gpu_lib, oracle_mod, group):
    """Later search passes of a scan start from the previous pass's neighbours (liinit_set_reseed, default on): bit-identical per-point state
    and accumulators to a search from scratch, equal to the oracle; a map update in between switches the seed off."""
    c = scenes.make_config("C2", N=20000, M=200000, open_air_frac=0.02)
    p = c["pose_init"]
    p2 = scenes.perturb_pose(p, 78, dtheta_deg=0.08, dpos=0.02)
    om = oracle_mod.OracleMap(c["ds"], _best_backend(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    Ho, bo, mo = osc.iterate(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, False, True)
    near_o = osc.get()["near_xyz"].copy()
    res = {}
    for seeded in (False, True):     # (the seeded round ends with a map update on both sides)
        g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_group_lanes=group)
        g.set_reseed(seeded)
        g.map_build(c["map_xyz"])
        g.scan_upload(c["body_xyz"])
        g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
        H, b, m, _ = g.icp_iterate(p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, False, True)
        res[seeded] = (H, b, m, g.scan_state(), g.last_pass_kernel_times()[0])
        assert m == mo and _relerr(H, Ho) <= 1e-9 and _relerr(b, bo) <= 1e-9
        assert np.array_equal(res[seeded][3]["near_xyz"], near_o)
        if seeded:
            extra = _world(c["body_xyz"][:3000], c["pose_gt"]) + np.float32(0.02)
            assert g.map_add_points(extra, True) == om.add_points(extra, True)
            H3, b3, m3, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
            Ho3, bo3, mo3 = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
            assert m3 == mo3 and _relerr(H3, Ho3) <= 1e-9 and np.array_equal(g.scan_state()["near_xyz"], osc.get()["near_xyz"])
        g.close()
    (Hs, bs, ms, sts, ts), (Hu, bu, mu, stu, tu) = res[True], res[False]
    assert ms == mu and np.array_equal(Hs, Hu) and np.array_equal(bs, bu)
    for k in ("world", "near_xyz", "near_cnt", "selected"):
        assert np.array_equal(sts[k], stu[k]), k
    print(f"seeded search kernel {ts:.4f} ms vs {tu:.4f} ms from scratch (G = {group or 'auto'})")
