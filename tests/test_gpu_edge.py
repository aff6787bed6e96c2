"""Edge cases of the C-ABI on the GPU: input layouts, empty / tiny / ragged inputs, invalid values, capacity and
call-order errors, large coordinates, slab growth, brick sizes, determinism, and the full BASELINE size."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu


def _world(body, p):
    return (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)


def _bk(orc):
    return 1 if orc.has_ikd() else 0


@pytest.fixture(scope="module")
def case():
    return scenes.make_config("C2", N=6000, M=60000, open_air_frac=0.02)


def test_point_layouts_equivalent(gpu_lib, case):
    c, p = case, case["pose_init"]
    res = []
    for stride in (3, 4, 12):
        def widen(a):
            out = np.zeros((len(a), stride), np.float32)
            out[:, :3] = a
            if stride == 12:      # pcl::PointXYZINormal: junk in the other fields must be ignored
                out[:, 3:] = 7.0
            return out
        g = gpu_lib.LiInitGpu(c["ds"], max_map_points=200000, max_scan_points=10000)
        g.map_build(widen(c["map_xyz"]))
        g.scan_upload(widen(c["body_xyz"]))
        res.append(g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True))
        g.close()
    for r in res[1:]:
        assert r[2] == res[0][2] and np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1])


def test_empty_map_and_tiny_scans(gpu_lib, case):
    c, p = case, case["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=100000, max_scan_points=10000)
    assert g.map_validnum() == 0
    g.scan_upload(c["body_xyz"][:100])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)   # search in an empty map
    assert m == 0 and not H.any() and not b.any() and rs == 0.0
    x, d2, cnt = g.nearest_search(np.zeros((5, 3), np.float32))
    assert not cnt.any() and np.all(d2 == -1)
    g.map_build(c["map_xyz"][:3])                                                       # fewer than 5 map points
    H, b, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    assert m == 0
    g.map_build(c["map_xyz"])
    for n in (1, 31, 33, 257):                                                          # ragged warp / block tails
        g.scan_upload(c["body_xyz"][:n])
        H, b, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True)
        st = g.scan_state()
        assert m == int(st["selected"].sum()) and len(st["selected"]) == n
    g.close()


def test_non_finite_and_far_points_are_ignored(gpu_lib, oracle_mod, case):
    c, p = case, case["pose_init"]
    body = c["body_xyz"][:2000].copy()
    clean = body.copy()
    bad = [3, 500, 1999]
    body[3] = [np.nan, 0, 0]
    body[500] = [np.inf, 1, 1]
    body[1999] = [1e30, 0, 0]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=200000, max_scan_points=10000)
    g.map_build(c["map_xyz"])
    g.scan_upload(body)
    H, b, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    st = g.scan_state()
    assert not st["selected"][bad].any() and np.isfinite(H).all() and np.isfinite(b).all()
    keep = np.ones(len(body), bool)
    keep[bad] = False
    g.scan_upload(clean[keep])
    H2, b2, m2, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    assert m2 == m and np.allclose(H2, H, rtol=1e-12) and np.allclose(b2, b, rtol=1e-10, atol=1e-12)
    # non-finite map points are dropped at insertion
    mp = c["map_xyz"][:1000].copy()
    mp[10] = [np.nan, np.nan, np.nan]
    g.map_build(mp)
    assert g.map_validnum() == 999
    g.close()


def test_errors_are_reported_not_swallowed(gpu_lib, case):
    c, p = case, case["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=1000, max_scan_points=500)
    with pytest.raises(gpu_lib.LiInitError) as e:
        g.scan_upload(c["body_xyz"][:501])
    assert e.value.code == -3
    with pytest.raises(gpu_lib.LiInitError) as e:
        g.map_build(c["map_xyz"][:1001])
    assert e.value.code == -3
    g.map_build(c["map_xyz"][:1000])
    g.scan_upload(c["body_xyz"][:100])
    with pytest.raises(gpu_lib.LiInitError) as e:      # reuse pass before any search pass
        g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, False)
    assert e.value.code == -1
    with pytest.raises(ValueError):
        g.map_build(np.zeros((10, 5), np.float32))       # unsupported stride
    g.close()
    with pytest.raises(gpu_lib.LiInitError):
        gpu_lib.LiInitGpu(-1.0)


def test_large_coordinates_keep_exactness(gpu_lib, oracle_mod):
    """Scene shifted 6 km from the origin: float cell assignment is coarse there; pruning margins must keep the kNN exact."""
    c = scenes.make_config("C2", N=4000, M=50000, open_air_frac=0.02)
    off = np.array([6000.0, -4500.0, 300.0])
    mp = (c["map_xyz"].astype(np.float64) + off).astype(np.float32)
    p = c["pose_init"]
    q = (_world(c["body_xyz"], p).astype(np.float64) + off).astype(np.float32)
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=200000, max_scan_points=10000)
    g.map_build(mp)
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(mp)
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.close()


@pytest.mark.parametrize("brick", [1, 2, 3, 4])
def test_brick_sizes(gpu_lib, oracle_mod, case, brick):
    c, p = case, case["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=200000, max_scan_points=10000, brick_cells_log2=brick)
    g.map_build(c["map_xyz"])
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    q = _world(c["body_xyz"], p)
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.close()


def test_slab_growth_over_many_batches(gpu_lib, oracle_mod):
    """Many insert batches into the same region force slab reallocation; the live set must keep matching."""
    ds = 0.15
    scene = scenes.box_scene(12.0, 9.0, 4.0)
    g = gpu_lib.LiInitGpu(ds, max_map_points=400000, max_scan_points=20000)
    om = oracle_mod.OracleMap(ds, _bk(oracle_mod))
    first = scenes.map_points(scene, 0.3, seed=3)
    g.map_build(first)
    om.build(first)
    rng = np.random.default_rng(9)
    for k in range(12):
        pts = scenes.map_points(scene, 0.11, seed=20 + k)
        pts = pts[rng.permutation(len(pts))[:8000]]
        down = bool(k % 3)
        g.map_add_points(pts, down)
        om.add_points(pts, down)
        assert g.map_validnum() == om.validnum()
    a = g.map_download()
    b = om.flatten()
    assert set(map(bytes, a)) == set(map(bytes, b))
    st = g.map_stats()
    assert st["pool_used"] <= st["pool_cap"]
    g.close()


def test_pass_is_bit_reproducible(gpu_lib, case):
    c, p = case, case["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=200000, max_scan_points=10000)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    outs = [g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True) for _ in range(3)]
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and o[2] == outs[0][2]
    g.close()


def test_full_baseline_size_properties(gpu_lib, oracle_mod):
    """BASELINE config 2 at full size (240k-point scan, 5M-point map): size-independent properties + a sampled
    comparison with the oracle (the verbatim ikd-Tree when oracle/_ref is present)."""
    c = scenes.make_config("C2")
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=6_000_000, max_scan_points=250_000)
    g.map_build(c["map_xyz"])
    assert g.map_validnum() == 5_000_000
    g.scan_upload(c["body_xyz"])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
    st = g.scan_state()
    assert m == int(st["selected"].sum()) and 0.9 * 240000 < m <= 240000
    assert np.array_equal(H, H.T) and np.all(np.linalg.eigvalsh(H[:6, :6]) > 0) and rs > 0
    full = st["near_cnt"] == 5
    assert not st["selected"][~full].any()
    d = np.linalg.norm(st["near_xyz"].astype(np.float64) - st["world"][:, None, :].astype(np.float64), axis=2)
    assert np.all(np.diff(d[full], axis=1) >= -1e-6) and np.all(d[full] ** 2 <= 5.0 + 1e-4)
    nv = st["normvec"][st["selected"].astype(bool)]
    assert np.allclose(np.linalg.norm(nv[:, :3], axis=1), 1.0, atol=1e-5)
    # linearity of the accumulators: two half scans add up to the whole
    half = len(c["body_xyz"]) // 2
    acc = np.zeros((12, 12))
    mm = 0
    for part in (c["body_xyz"][:half], c["body_xyz"][half:]):
        g.scan_upload(part)
        Hp, bp, mp_, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
        acc += Hp
        mm += mp_
    assert mm == m and np.allclose(acc, H, rtol=1e-11)
    # sampled exact comparison with the oracle
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    idx = np.random.default_rng(1).permutation(len(c["body_xyz"]))[:20000]
    ox, od, oc, _ = om.knn(st["world"][idx])
    assert np.array_equal(oc, st["near_cnt"][idx]) and np.array_equal(ox, st["near_xyz"][idx])
    g.close()


def test_downsample_box_membership_is_geometric(gpu_lib, oracle_mod):
    """The float boxes [fl(k ds), fl(fl(k ds) + ds)) of neighbouring k overlap by one ulp: an existing lattice point belongs to two boxes
    although it is filed under one voxel id; Add_Points(downsample) must find it from either (ikd_Tree.cpp:633,980). Found by the
    CPU fuzz (tools/emul_fuzz.py); same scenario as tests/test_liinit_emul.py::test_emul_downsample_box_membership_is_geometric."""
    def lattice(rng, n, ds, off):
        k = rng.integers(-40, 40, (n, 3)).astype(np.float32)
        a = (k * np.float32(ds)).astype(np.float32)
        j = rng.integers(0, 3, n)
        a = np.where(j[:, None] == 0, a, np.where(j[:, None] == 1, np.nextafter(a, np.float32(1e9)), np.nextafter(a, np.float32(-1e9))))
        return (a.astype(np.float64) + off).astype(np.float32)

    def single_box(pts, ds):
        f, d = np.float32, np.float32(ds)
        c = np.floor(pts / d).astype(np.float32)
        bad = np.zeros(len(pts), bool)
        for a in range(3):
            for k in (-1, 0, 1):
                mn = ((c[:, a] + f(k)) * d).astype(f)
                inside = (pts[:, a] >= mn) & (pts[:, a] < (mn + d).astype(f))
                bad |= inside if k != 0 else ~inside
        return np.ascontiguousarray(pts[~bad])

    ds = 0.15
    for off in (0.0, 900.0):
        rng = np.random.default_rng(1011)
        o = off * np.array([1.0, -0.7, 0.1])
        first = lattice(rng, 3500, ds, o)
        g = gpu_lib.LiInitGpu(ds, max_map_points=60000, max_scan_points=4000)
        om = oracle_mod.OracleMap(ds, _bk(oracle_mod))
        g.map_build(first)
        om.build(first)
        for k in range(3):
            pts = single_box(lattice(rng, 1500, ds, o), ds)
            assert g.map_add_points(pts, True) == om.add_points(pts, True)
            assert g.map_validnum() == om.validnum()
        assert set(map(bytes, g.map_download())) == set(map(bytes, om.flatten()))
        g.close()


@pytest.mark.parametrize("group", [0, 4, 16, 32])
def test_hollow_with_dense_surroundings(gpu_lib, group):
    """tests/hollow_case.py: a shell that finds 44 non-empty bricks in one probing round (more than the 32 entries a 32-lane group's
    list once had: the surplus went into the next warp's list). 4000 queries so that every warp of a block is at work at the same time."""
    import hollow_case as hc
    mp, qs = hc.hollow_map_and_queries(4000)
    g = gpu_lib.LiInitGpu(hc.DS, max_map_points=80000, max_scan_points=len(qs) + 16, knn_group_lanes=group)
    g.map_build(mp)
    live = g.map_download()
    g.scan_upload(qs)
    I, z = np.eye(3), np.zeros(3)
    g.icp_iterate(I, z, I, z, False, True)
    st = g.scan_state()
    want = hc.brute_force_sets(live, qs[:400])
    for i in range(400):
        assert st["near_cnt"][i] == 5 and set(map(bytes, st["near_xyz"][i])) == want[i], (group, i)
    # every query, cheaply: the five distances against a brute-force 5th-neighbour distance
    d5 = np.array([np.sort(((live - q) ** 2).astype(np.float32).sum(1))[4] for q in qs[::10]])
    got = ((st["near_xyz"][::10] - qs[::10, None, :]) ** 2).astype(np.float32).sum(2).max(1)
    assert np.allclose(got, d5, rtol=1e-5)
    g.close()


def test_pool_full_leaves_a_consistent_map(gpu_lib):
    """A batch that exhausts the point pool fails once with LIINIT_ERR_CAPACITY; what is stored stays intact under the concurrent
    reservations of the failing batch (no slab handed out twice), the allocator is back inside the pool, later batches fit again.
    (Same scenario as tests/test_liinit_emul.py::test_emul_pool_full_leaves_a_consistent_map, here with thousands of warps reserving at once.)"""
    ds = 0.15
    rng = np.random.default_rng(11)
    g = gpu_lib.LiInitGpu(ds, max_map_points=20000, max_scan_points=100, hash_capacity_log2=19)   # hash large enough: the POOL runs out
    cap = g.map_stats()["pool_cap"]
    first = rng.uniform(-20, 20, (15000, 3)).astype(np.float32)
    g.map_build(first)
    assert g.map_validnum() == 15000
    used0 = g.map_stats()["pool_used"]
    far = rng.uniform(-2000, 2000, (cap // 8, 3)).astype(np.float32)      # one new brick (a 16-point slab) per point: twice the pool
    with pytest.raises(gpu_lib.LiInitError) as e:
        g.map_add_points(far, False)
    assert e.value.code == -3
    st = g.map_stats()
    assert used0 <= st["pool_used"] <= cap
    live = g.map_download()
    have = set(map(bytes, live))
    assert len(live) == g.map_validnum() == len(have)
    assert set(map(bytes, first)) <= have and have <= set(map(bytes, first)) | set(map(bytes, far))
    g.map_delete_boxes(np.array([[-3000, -3000, -3000, -21, 3000, 3000], [21, -3000, -3000, 3000, 3000, 3000]], np.float32))
    g.map_compact()
    more = rng.uniform(-20, 20, (2000, 3)).astype(np.float32)
    assert g.map_add_points(more, False) == 2000
    live2 = set(map(bytes, g.map_download()))
    assert set(map(bytes, first)) | set(map(bytes, more)) <= live2 and len(live2) == g.map_validnum()
    x, d, cnt = g.nearest_search(first[:500])
    assert np.all(d[:, 0] == 0) and np.array_equal(x[:, 0], first[:500])
    g.close()


def test_coupled_boxes_are_walked_in_batch_order(gpu_lib, oracle_mod):
    """tests/coupled_case.py: downsample boxes that share an existing point (one-ulp overlap of their float boxes) and both receive new
    points in one batch -- changed-box count and live set equal the reference's sequential walk for every batch order."""
    import coupled_case as cc
    f = np.float32
    base, batches = cc.scene()
    counts = []
    for order, batch in enumerate(batches):
        g = gpu_lib.LiInitGpu(cc.DS, max_map_points=20000, max_scan_points=100, hash_capacity_log2=12)
        om = oracle_mod.OracleMap(cc.DS, 0)            # the restated tree: deterministic counters
        g.map_build(base)
        om.build(base)
        a, b = g.map_add_points(batch, True), om.add_points(batch, True)
        assert a == b, (order, a, b)
        counts.append(a)
        live, want = g.map_download(), om.flatten()
        assert g.map_validnum() == om.validnum() and set(map(bytes, live)) == set(map(bytes, np.ascontiguousarray(want, np.float32))), order
        more = (batch + f(0.004)).astype(f)
        assert g.map_add_points(more, True) == om.add_points(more, True)
        assert set(map(bytes, g.map_download())) == set(map(bytes, np.ascontiguousarray(om.flatten(), np.float32))), order
        g.close()
    assert counts[0] != counts[1]
