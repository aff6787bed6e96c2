"""Randomised differential test of the ICP pass on the CPU: random surface maps, scans, poses, 6 / 12 columns, both indexes -- search
pass, reuse pass at a moved pose, map_incremental, a search on the updated map -- through the emulated library against the oracle
(verbatim ikd-Tree + restated loop). usage: python tools/emul_fuzz_pass.py [n_scenarios] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import liinit_emul as le
from oracle import oracle as orc
from lidar_imu_init_b200 import scenes

n_sc = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bk = 1 if orc.has_ikd() else 0
rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
t0 = time.time()
for sc in range(n_sc):
    rng = np.random.default_rng(seed0 * 7919 + sc)
    imu_en = bool(rng.integers(0, 2))
    index = int(rng.integers(1, 3))   # LIINIT_KNN_BRICKS, LIINIT_KNN_CELLS
    N, M = int(rng.integers(300, 2500)), int(rng.integers(8000, 50000))
    c = scenes.make_config("C2", seed=int(rng.integers(1, 1000)), N=N, M=M, open_air_frac=float(rng.choice([0.0, 0.02, 0.2])), imu_en=imu_en,
                           order=str(rng.choice(["voxel", "shuffle", "morton"])))
    if rng.integers(0, 4) == 0:
        # a volumetric map with hollows instead of surfaces (dense surroundings, empty neighbourhoods: the shells of the search grow to
        # their last size and find many bricks at once), scan points inside and around the hollows
        vol = rng.uniform(-4.5, 4.5, (M, 3))
        centres = rng.uniform(-3, 3, (int(rng.integers(2, 8)), 3))
        for cc in centres:
            vol = vol[((vol - cc) ** 2).sum(1) > rng.uniform(0.6, 1.6) ** 2]
        c = dict(c)
        gt = c["pose_gt"]
        c["map_xyz"] = np.ascontiguousarray(vol, np.float32)
        w = np.concatenate([centres[rng.integers(0, len(centres), N // 2)] + rng.normal(0, 0.25, (N // 2, 3)), rng.uniform(-4, 4, (N - N // 2, 3))])
        # body points such that the ground-truth pose maps them onto w
        c["body_xyz"] = np.ascontiguousarray((gt.R_LI.T @ (gt.rot_end.T @ (w - gt.pos_end).T - gt.T_LI[:, None])).T, np.float32)
    p = scenes.perturb_pose(c["pose_gt"], int(rng.integers(0, 10**6)), dtheta_deg=float(rng.choice([0.0, 0.1, 0.5, 3.0])), dpos=float(rng.choice([0.0, 0.05, 0.5])))
    g = le.EmulGpu(c["ds"], max_map_points=3 * M + 20000, max_scan_points=N + 10, knn_index=index, knn_group_lanes=int(rng.choice([0, 2, 4, 8, 16, 32])),
                   knn_seed_radius_cells=float(rng.choice([0.0, 1.0, 4.0])), hash_capacity_log2=14)
    om = orc.OracleMap(c["ds"], bk)
    g.map_build(c["map_xyz"]); om.build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    osc = orc.OracleScan(c["body_xyz"])
    tag = f"scenario {sc}: index {index} imu {int(imu_en)} N {N} M {M}"
    for it in range(2):
        H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
        Ho, bo, mo = osc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
        st, so = g.scan_state(), osc.get()
        assert m == mo, (tag, "m", m, mo)
        for k in ("world", "near_cnt", "near_xyz", "selected"):
            assert np.array_equal(st[k], so[k]), (tag, k)
        sel = so["selected"].astype(bool)
        assert np.array_equal(st["normvec"][sel], so["normvec"][sel]), (tag, "normvec")
        if mo > 0:
            assert rel(H, Ho) <= 1e-9 and rel(b, bo) <= 1e-9, (tag, "H")
        # a later search pass of the same scan, map untouched: started from the stored neighbours (seeded), must equal a search from scratch
        p1 = scenes.perturb_pose(p, int(rng.integers(0, 10**6)), dtheta_deg=float(rng.choice([0.02, 0.3])), dpos=float(rng.choice([0.005, 0.08])))
        H1, b1, m1, _ = g.icp_iterate(p1.rot_end, p1.pos_end, p1.R_LI, p1.T_LI, imu_en, True)
        Ho1, bo1, mo1 = osc.iterate(om, p1.rot_end, p1.pos_end, p1.R_LI, p1.T_LI, imu_en, True)
        st, so = g.scan_state(), osc.get()
        assert m1 == mo1 and (mo1 == 0 or (rel(H1, Ho1) <= 1e-9 and rel(b1, bo1) <= 1e-9)), (tag, "seeded pass")
        for k in ("world", "near_cnt", "selected"):
            assert np.array_equal(st[k], so[k]), (tag, "seeded pass", k)
        found = np.arange(5)[None, :] < so["near_cnt"][:, None]      # (ranks beyond the count hold whatever an earlier pass left there)
        assert np.array_equal(st["near_xyz"][found], so["near_xyz"][found]), (tag, "seeded pass", "near_xyz")
        p = p1
        p2 = scenes.perturb_pose(p, int(rng.integers(0, 10**6)), dtheta_deg=0.05, dpos=0.01)
        H2, b2, m2, _ = g.icp_iterate(p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
        Ho2, bo2, mo2 = osc.iterate(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
        assert m2 == mo2 and (mo2 == 0 or (rel(H2, Ho2) <= 1e-9 and rel(b2, bo2) <= 1e-9)), (tag, "reuse")
        assert np.array_equal(g.scan_state()["selected"], osc.get()["selected"]), (tag, "reuse flags")
        na, nn = g.map_incremental(p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, c["ds"])
        _, oa, on, _ = osc.map_incremental(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, c["ds"])
        assert (na, nn) == (oa, on) and g.map_validnum() == om.validnum(), (tag, "map_incremental", na, nn, oa, on)
        p = p2
    assert set(map(bytes, g.map_download())) == set(map(bytes, om.flatten())), (tag, "live set")
    g.close()
    print(tag, "| ok", flush=True)
print(f"emul_fuzz_pass: {n_sc} scenarios, no mismatch ({time.time() - t0:.0f} s)")
