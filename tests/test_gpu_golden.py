"""GPU vs the committed golden vectors (tests/golden/*.npz, generated with the reference's verbatim ikd-Tree).
Does not touch the oracle: the fixtures alone pin the CUDA path."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "c*.npz")))   # ICP-path fixtures (li_init_log.npz: test_calib.py)


def _pose(a):
    return a[0:9].reshape(3, 3), a[9:12], a[12:21].reshape(3, 3), a[21:24]


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(g)[:-4] for g in GOLD])
@pytest.mark.parametrize("brick", [2, 3])
def test_gpu_reproduces_golden(gpu_lib, path, brick):
    G = np.load(path)
    imu_en = bool(G["imu_en"])
    ds = float(G["ds"])
    g = gpu_lib.LiInitGpu(ds, max_map_points=200000, max_scan_points=10000, brick_cells_log2=brick)
    g.map_build(G["map_xyz"])
    g.scan_upload(G["body_xyz"])
    H, b, m, _ = g.icp_iterate(*_pose(G["pose_init"]), imu_en, True)
    st = g.scan_state()
    assert m == int(G["s_m"])
    assert np.array_equal(st["world"], G["s_world"])
    assert np.array_equal(st["near_cnt"], G["s_near_cnt"])
    assert np.array_equal(st["near_xyz"], G["s_near_xyz"])
    assert np.array_equal(st["selected"], G["s_selected"])
    sel = G["s_selected"].astype(bool)
    assert np.array_equal(st["normvec"][sel], G["s_normvec"][sel])      # f32 normal + residual: bit-equal
    assert _rel(H, G["s_HtH"]) <= 1e-9 and _rel(b, G["s_Htr"]) <= 1e-9   # f64 accumulators: tolerance 1e-9 relative
    H2, b2, m2, _ = g.icp_iterate(*_pose(G["pose_2"]), imu_en, False)
    assert m2 == int(G["r_m"]) and _rel(H2, G["r_HtH"]) <= 1e-9 and _rel(b2, G["r_Htr"]) <= 1e-9
    st2 = g.scan_state()
    assert np.array_equal(st2["selected"], G["r_selected"])
    na, nn = g.map_incremental(*_pose(G["pose_gt"]), ds)
    assert (na, nn) == (int(G["mi_n_add"]), int(G["mi_n_nod"]))
    live = g.map_download()
    live = live[np.lexsort((live[:, 2], live[:, 1], live[:, 0]))]
    assert np.array_equal(live, G["mi_live"])
    g.close()
