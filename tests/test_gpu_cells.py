"""Both spatial indexes of the 5-NN kernel, selected explicitly (liinit_config.knn_index): LIINIT_KNN_BRICKS = lockstep
lane groups over whole bricks (knn_kernels.cuh), LIINIT_KNN_CELLS = per-brick cell directory, one scan point per thread
(cells.cuh). Same bars as test_gpu_parity.py: identical neighbours / flags / f32 normals vs the oracle, HtH / Htr 1e-9;
plus directory maintenance under every kind of map update (plain insert, downsample insert, box delete, slab growth)."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu

BRICKS, CELLS = 1, 2
REL = 1e-9


def _world(body, p):
    return (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)


def _bk(orc):
    return 1 if orc.has_ikd() else 0


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def small_case():
    return scenes.make_config("C2", N=20000, M=200000, open_air_frac=0.02)


@pytest.mark.parametrize("index", [BRICKS, CELLS])
@pytest.mark.parametrize("rho_cells", [0.0, 1.0, 5.0])
def test_index_knn_matches_oracle(gpu_lib, oracle_mod, small_case, index, rho_cells):
    c = small_case
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_index=index, knn_seed_radius_cells=rho_cells)
    assert g.knn_index() == index
    g.map_build(c["map_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    for pose in ("pose_init", "pose_gt"):
        q = _world(c["body_xyz"], c[pose])
        gx, gd, gc = g.nearest_search(q)
        ox, od, oc, _ = om.knn(q)
        assert np.array_equal(gc, oc)
        assert np.array_equal(gd, od), f"d2 mismatch at {np.argwhere(gd != od)[:5]}"
        assert np.array_equal(gx, ox)
    g.close()


@pytest.mark.parametrize("search", [1, 2, 3])
def test_cells_all_loop_shapes(gpu_lib, oracle_mod, small_case, search, monkeypatch):
    """The three searches over the directory (shells on cells, growing boxes, enumerate + stream) are kept
    selectable for A/B (LIINIT_CELLS_SEARCH, read when the context is created): all of them must stay exact."""
    monkeypatch.setenv("LIINIT_CELLS_SEARCH", str(search))
    c = small_case
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_index=CELLS)
    g.map_build(c["map_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    q = _world(c["body_xyz"], c["pose_init"])
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.scan_upload(c["body_xyz"][:1000])          # ragged tail: 1000 = 7 blocks of 128 + 104 (idle lanes must keep voting)
    p = c["pose_init"]
    H, b, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    osc = oracle_mod.OracleScan(c["body_xyz"][:1000])
    Ho, bo, mo = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    assert m == mo and _relerr(H, Ho) <= REL and _relerr(b, bo) <= REL
    g.close()


@pytest.mark.parametrize("imu_en", [False, True])
def test_cells_search_and_reuse_pass(gpu_lib, oracle_mod, imu_en):
    c = scenes.make_config("C2", N=20000, M=200000, open_air_frac=0.02, imu_en=imu_en)
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_index=CELLS)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    Ho, bo, mo = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    assert m == mo and m > 15000
    st, so = g.scan_state(), osc.get()
    assert np.array_equal(st["world"], so["world"])
    assert np.array_equal(st["near_cnt"], so["near_cnt"])
    assert np.array_equal(st["near_xyz"], so["near_xyz"])
    assert np.array_equal(st["selected"], so["selected"])
    sel = so["selected"].astype(bool)
    assert np.array_equal(st["normvec"][sel], so["normvec"][sel])
    assert _relerr(H, Ho) <= REL and _relerr(b, bo) <= REL
    p2 = scenes.perturb_pose(p, 77, dtheta_deg=0.05, dpos=0.01)
    H2, b2, m2, _ = g.icp_iterate(p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    Ho2, bo2, mo2 = osc.iterate(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    assert m2 == mo2 and _relerr(H2, Ho2) <= REL and _relerr(b2, bo2) <= REL
    g.close()


def test_both_indexes_give_the_same_pass_bit_for_bit(gpu_lib, small_case):
    """Same neighbours in the same (ascending) order -> the plane / Jacobian / reduction kernel sees identical input."""
    c, p = small_case, small_case["pose_init"]
    outs = []
    for index in (BRICKS, CELLS):
        g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=50000, knn_index=index)
        g.map_build(c["map_xyz"])
        g.scan_upload(c["body_xyz"])
        outs.append((g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True), g.scan_state()))
        g.close()
    (ra, sa), (rb, sb) = outs
    assert ra[2] == rb[2] and np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k


def test_cells_directory_follows_map_updates(gpu_lib, oracle_mod):
    """map_incremental (downsample insert + plain insert), extra batches, a box delete: after each, searches on the
    re-sorted slabs must still agree with the oracle's tree."""
    c = scenes.make_config("C2", N=20000, M=150000, open_air_frac=0.02)
    p, gt = c["pose_init"], c["pose_gt"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=600000, max_scan_points=50000, knn_index=CELLS)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    osc = oracle_mod.OracleScan(c["body_xyz"])
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    na, nn = g.map_incremental(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"])
    _, oa, on, _ = osc.map_incremental(om, gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"])
    assert (na, nn) == (oa, on) and g.map_validnum() == om.validnum()

    def same_search(q):
        gx, gd, gc = g.nearest_search(q)
        ox, od, oc, _ = om.knn(q)
        assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)

    q = _world(c["body_xyz"][:6000], gt)
    same_search(q)
    new = _world(c["body_xyz"], gt) + np.float32(0.013)
    for lo, hi, down in ((0, 7000, True), (7000, 12000, False), (5000, 16000, True)):
        g.map_add_points(new[lo:hi], down)
        om.add_points(new[lo:hi], down)
        assert g.map_validnum() == om.validnum()
        same_search(q)
    sc = c["scene"]
    boxes = np.array([[-1, -1, -1, 0.4 * sc.L, sc.W + 1, sc.H + 1]], np.float32)
    assert g.map_delete_boxes(boxes) == om.delete_boxes(boxes)
    same_search(q)
    assert set(map(bytes, g.map_download())) == set(map(bytes, om.flatten()))
    # a second scan runs the whole pass on the updated map
    g.scan_upload(c["body_xyz"])
    H, b, m, _ = g.icp_iterate(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, False, True)
    Ho, bo, mo = osc.iterate(om, gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, False, True)
    assert m == mo and _relerr(H, Ho) <= REL and _relerr(b, bo) <= REL
    g.close()


def test_cells_slab_growth_and_rebuild(gpu_lib, oracle_mod):
    ds = 0.15
    scene = scenes.box_scene(12.0, 9.0, 4.0)
    g = gpu_lib.LiInitGpu(ds, max_map_points=400000, max_scan_points=20000, knn_index=CELLS)
    om = oracle_mod.OracleMap(ds, _bk(oracle_mod))
    first = scenes.map_points(scene, 0.3, seed=3)
    g.map_build(first)
    om.build(first)
    rng = np.random.default_rng(9)
    q = (scenes.map_points(scene, 0.2, seed=77)[:4000] + rng.normal(0, 0.05, (4000, 3))).astype(np.float32)
    for k in range(6):
        pts = scenes.map_points(scene, 0.11, seed=20 + k)
        pts = pts[rng.permutation(len(pts))[:8000]]
        g.map_add_points(pts, bool(k % 3))
        om.add_points(pts, bool(k % 3))
        gx, gd, gc = g.nearest_search(q)
        ox, od, oc, _ = om.knn(q)
        assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.map_build(first)          # Build replaces the map: stale directory entries / super-brick bits must not leak
    om2 = oracle_mod.OracleMap(ds, _bk(oracle_mod))
    om2.build(first)
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om2.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.close()


def test_cells_dense_cloud_oversized_brick_and_large_coordinates(gpu_lib, oracle_mod):
    rng = np.random.default_rng(11)
    off = np.array([6000.0, -4500.0, 300.0])
    core = rng.uniform(9.7, 10.7, size=(66000, 3))       # > 0xfff0 points in one brick: searched as a whole slab
    dense = rng.uniform(8.0, 13.0, size=(60000, 3))
    sparse = rng.uniform(0.0, 20.0, size=(30000, 3))
    mp = (np.concatenate([core, dense, sparse]) + off).astype(np.float32)
    g = gpu_lib.LiInitGpu(0.15, max_map_points=400000, max_scan_points=20000, knn_index=CELLS)
    g.map_build(mp)
    q = (np.concatenate([rng.uniform(8.0, 13.0, size=(1500, 3)), rng.uniform(0.0, 20.0, size=(1500, 3)),
                         rng.uniform(30.0, 40.0, size=(50, 3))]) + off).astype(np.float32)
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = oracle_mod.knn_bruteforce(mp, q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od)
    same = np.all(np.diff(od, axis=1) != 0, axis=1)
    assert np.array_equal(gx[same], ox[same])
    g.close()


def test_cells_attach_host_matches_upload(gpu_lib):
    import torch
    c = scenes.make_config("C2", N=6000, M=60000, open_air_frac=0.02)
    p, body = c["pose_init"], c["body_xyz"]
    n = len(body)
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=200000, max_scan_points=n + 16, knn_index=CELLS)
    g.map_build(c["map_xyz"])
    g.scan_upload(body)
    ref = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True)
    ref_state = g.scan_state()
    for stride in (3, 4, 12):
        host = torch.full((n, stride), 7.0, dtype=torch.float32).pin_memory()
        host[:, :3] = torch.from_numpy(body)
        g.scan_attach_ptr(host.data_ptr(), stride, n)
        got = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True)
        for a, b in zip(ref, got):
            assert np.array_equal(np.asarray(a), np.asarray(b))
        st = g.scan_state()
        for k in ref_state:
            assert np.array_equal(ref_state[k], st[k]), k
        assert np.array_equal(g.scan_body(), body)
    g.close()


def test_cells_needs_default_brick_size(gpu_lib, small_case):
    """The directory is defined for 8x8x8-voxel bricks; another brick size keeps the brick search (documented)."""
    g = gpu_lib.LiInitGpu(small_case["ds"], max_map_points=100000, max_scan_points=1000, knn_index=CELLS, brick_cells_log2=2)
    assert g.knn_index() == BRICKS
    g.close()
    with pytest.raises(gpu_lib.LiInitError):
        gpu_lib.LiInitGpu(small_case["ds"], max_map_points=100000, max_scan_points=1000, knn_index=7)
