"""The cell-directory 5-NN (lidar_imu_init_b200/csrc/cells.cuh) checked on the CPU: the SAME LI_HD source the sm_100a
kernels are built from is compiled for the host by tests/emul/cells_emul.cpp and compared with the oracle (verbatim
ikd-Tree when oracle/_ref is present) and with brute force. Bit-exact: counts, squared distances, neighbour points."""
import numpy as np
import pytest

import cells_emul as ce
from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.skipif(not ce.available(), reason="g++ or the CUDA vector-type headers are missing")


def _world(body, p):
    return (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)


def _bk(orc):
    return 1 if orc.has_ikd() else 0


@pytest.fixture(scope="module")
def case():
    return scenes.make_config("C2", N=20000, M=200000, open_air_frac=0.02)


VARIANTS = [1, 2, 3]   # 1 = shells on cells (knn5_cells), 2 = growing boxes (knn5_boxes), 3 = enumerate + stream (knn5_stream); 0 = the kernels' default


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("rho", [0.15, 0.3, 1.0])
def test_cells_knn_matches_oracle(oracle_mod, case, rho, variant):
    c = case
    E = ce.CellsEmul(c["map_xyz"], c["ds"])
    assert E.check_directory() == 0
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(c["map_xyz"])
    for pose in ("pose_init", "pose_gt"):
        q = _world(c["body_xyz"], c[pose])
        gx, gd, gc, st = E.knn(q, rho=rho, stats=True, variant=variant)
        ox, od, oc, _ = om.knn(q)
        assert np.array_equal(gc, oc)
        assert np.array_equal(gd, od), f"d2 mismatch at {np.argwhere(gd != od)[:5]}"
        assert np.array_equal(gx, ox)
        assert (gc == 0).sum() >= 300 and (gc == 5).sum() > 15000
        # the point of the directory: far fewer candidates than a whole-brick scan (~64 points per brick crossing)
        if rho <= 0.3:
            assert st[:, 6].mean() < 120
    E.close()


@pytest.mark.parametrize("variant", VARIANTS + [0])
def test_cells_large_coordinates(oracle_mod, variant):
    """6 km from the origin the float box indices are coarse; margins and range slack must keep the search exact."""
    c = scenes.make_config("C2", N=4000, M=50000, open_air_frac=0.02)
    off = np.array([6000.0, -4500.0, 300.0])
    mp = (c["map_xyz"].astype(np.float64) + off).astype(np.float32)
    q = (_world(c["body_xyz"], c["pose_init"]).astype(np.float64) + off).astype(np.float32)
    E = ce.CellsEmul(mp, c["ds"])
    assert E.check_directory() == 0
    om = oracle_mod.OracleMap(c["ds"], _bk(oracle_mod))
    om.build(mp)
    gx, gd, gc = E.knn(q, variant=variant)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    E.close()


@pytest.mark.parametrize("variant", VARIANTS + [0])
def test_cells_directory_survives_appends(oracle_mod, case, variant):
    """Appended points land unsorted behind the sorted slab (k_ins_append); the refresh must re-sort and re-index."""
    c = case
    mp = c["map_xyz"]
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(mp))
    E = ce.CellsEmul(mp[perm[:80000]], c["ds"], hash_log2=17)
    for lo, hi in ((80000, 120000), (120000, 199000), (199000, len(mp))):
        E.add(mp[perm[lo:hi]])
        assert E.check_directory() == 0
    q = _world(c["body_xyz"][:8000], c["pose_init"])
    gx, gd, gc = E.knn(q, variant=variant)
    ox, od, oc, _ = oracle_mod.knn_bruteforce(mp, q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    E.close()


@pytest.mark.parametrize("variant", VARIANTS + [0])
def test_cells_dense_cloud_and_oversized_brick(oracle_mod, variant):
    """Build without downsampling accepts any density: many points per voxel, and one brick beyond the u16 directory
    (> 0xfff0 points) that must be searched as a whole slab."""
    rng = np.random.default_rng(11)
    dense = rng.uniform(10.05, 10.95, size=(70000, 3))          # inside one 1.2 m brick: [9.6, 10.8) or [10.8, 12.0) per axis -> up to 8 bricks
    core = rng.uniform(9.7, 10.7, size=(66000, 3))               # all in the brick [9.6, 10.8)^3
    sparse = rng.uniform(0.0, 20.0, size=(30000, 3))
    mp = np.concatenate([dense, core, sparse]).astype(np.float32)
    E = ce.CellsEmul(mp, 0.15, hash_log2=16)
    assert E.check_directory() == 0
    q = np.concatenate([rng.uniform(8.0, 13.0, size=(1500, 3)), rng.uniform(0.0, 20.0, size=(1500, 3)),
                        rng.uniform(30.0, 40.0, size=(50, 3))]).astype(np.float32)
    gx, gd, gc = E.knn(q, variant=variant)
    ox, od, oc, _ = oracle_mod.knn_bruteforce(mp, q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od)
    same = np.all(np.diff(od, axis=1) != 0, axis=1)             # exact distance ties may pick either point
    assert np.array_equal(gx[same], ox[same])
    assert (gc == 0).sum() >= 50
    E.close()


@pytest.mark.parametrize("variant", VARIANTS + [0])
def test_cells_grid_aligned_scene_distances(oracle_mod, variant):
    """BASELINE config 1 (noise-free planar grid): exact distance ties everywhere, so compare counts and distances."""
    c = scenes.make_config("C1")
    E = ce.CellsEmul(c["map_xyz"], c["ds"])
    assert E.check_directory() == 0
    q = _world(c["body_xyz"], c["pose_init"])
    gx, gd, gc = E.knn(q, variant=variant)
    ox, od, oc, _ = oracle_mod.knn_bruteforce(c["map_xyz"], q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od)
    d = ((gx.astype(np.float32) - q[:, None, :]) ** 2)
    E.close()


@pytest.mark.parametrize("variant", VARIANTS + [0])
def test_cells_rejects_non_finite_and_far_queries(case, variant):
    E = ce.CellsEmul(case["map_xyz"][:5000], case["ds"])
    q = np.array([[np.nan, 0, 0], [np.inf, 1, 1], [1e30, 0, 0], [0, -np.inf, 0]], np.float32)
    gx, gd, gc = E.knn(q, variant=variant)
    assert not gc.any() and np.all(gd == -1)
    E.close()


@pytest.mark.parametrize("variant", VARIANTS + [0])
def test_cells_empty_and_tiny_maps(oracle_mod, variant):
    E = ce.CellsEmul(np.zeros((0, 3), np.float32), 0.15)
    gx, gd, gc = E.knn(np.zeros((4, 3), np.float32), variant=variant)
    assert not gc.any()
    E.close()
    mp = np.array([[0.01, 0.02, 0.03], [0.5, 0.5, 0.5], [-0.2, 0.1, 2.0]], np.float32)
    E = ce.CellsEmul(mp, 0.15)
    q = np.array([[0, 0, 0], [0.4, 0.4, 0.4], [5, 5, 5]], np.float32)
    gx, gd, gc = E.knn(q, variant=variant)
    ox, od, oc, _ = oracle_mod.knn_bruteforce(mp, q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    E.close()
