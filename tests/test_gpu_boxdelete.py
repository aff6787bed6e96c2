"""Row N1: KD_TREE::Delete_Point_Boxes on the device map + the FoV segment logic that feeds it."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes

pytestmark = pytest.mark.gpu


def test_delete_boxes_matches_oracle(gpu_lib, oracle_mod):
    c = scenes.make_config("C2", N=5000, M=120000, open_air_frac=0.0)
    sc = c["scene"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=300000, max_scan_points=10000)
    g.map_build(c["map_xyz"])
    om = oracle_mod.OracleMap(c["ds"], 1 if oracle_mod.has_ikd() else 0)
    om.build(c["map_xyz"])
    boxes = np.array([[-1, -1, -1, 0.3 * sc.L, sc.W + 1, sc.H + 1],          # a slab of the building
                      [0.6 * sc.L, 0.5 * sc.W, -1, 0.8 * sc.L, sc.W + 1, 1.0]], np.float32)
    dg = g.map_delete_boxes(boxes)
    do = om.delete_boxes(boxes)
    assert dg == do > 1000 and g.map_validnum() == om.validnum()
    assert set(map(bytes, g.map_download())) == set(map(bytes, om.flatten()))
    # the map keeps working: searches and inserts after the delete agree with the oracle
    p = c["pose_init"]
    q = (p.rot_end @ (p.R_LI @ c["body_xyz"].T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)
    gx, gd, gc = g.nearest_search(q)
    ox, od, oc, _ = om.knn(q)
    assert np.array_equal(gc, oc) and np.array_equal(gd, od) and np.array_equal(gx, ox)
    g.map_add_points(q, True)
    om.add_points(q, True)
    assert g.map_validnum() == om.validnum()
    assert g.map_delete_boxes(np.zeros((0, 6), np.float32)) == 0
    g.close()


def test_fov_segment_bounds_map_growth(gpu_lib):
    """Drive the sensor through a long corridor; with the FoV boxes wired to the delete the live map stays bounded."""
    from lidar_imu_init_b200 import _build, host
    _build.build_host()
    ds, det, cube = 0.5, 20.0, 80.0
    g = gpu_lib.LiInitGpu(ds, max_map_points=400000, max_scan_points=20000)
    seg = host.FovSegmenter(cube, det)
    rng = np.random.default_rng(2)
    peak = 0
    for k in range(60):
        x0 = 4.0 * k
        pts = np.stack([x0 + rng.uniform(-det, det, 6000), rng.uniform(-8, 8, 6000), np.zeros(6000)], 1).astype(np.float32)
        if k == 0:
            g.map_build(pts)
        else:
            g.map_add_points(pts, True)
        boxes = seg.update([x0, 0.0, 0.0])
        if len(boxes):
            g.map_delete_boxes(boxes)
        peak = max(peak, g.map_validnum())
        live = g.map_download()
        assert live[:, 0].min() >= seg.box[0] - 1e-3 or k == 0
    assert peak < 16 * 2 * (cube + 2 * det) / (ds * ds) * 0.5   # bounded by the cube footprint, not by the 240 m driven
    g.close()
