"""Generate tests/golden/*.npz: inputs + outputs of the hot path computed HERE with the reference's verbatim
ikd-Tree (oracle/_ref, backend 1) and the restated measurement model. The fixtures travel to the GPU box,
where /root/reference does not exist; tests compare both the oracle and the CUDA path against them.

Run from the repo root in the build container:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_imu_init_b200 import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def world(body, p):
    return (p.rot_end @ (p.R_LI @ body.T.astype(np.float64) + p.T_LI[:, None]) + p.pos_end[:, None]).T.astype(np.float32)


def pose_arr(p):
    return np.concatenate([p.rot_end.reshape(9), p.pos_end, p.R_LI.reshape(9), p.T_LI])


def one(name, cfg, imu_en):
    assert orc.has_ikd(), "golden vectors must come from the verbatim ikd-Tree build (make -C oracle ref)"
    c = cfg
    p = c["pose_init"]
    om = orc.OracleMap(c["ds"], 1)
    om.build(c["map_xyz"])
    sc = orc.OracleScan(c["body_xyz"])
    out = dict(map_xyz=c["map_xyz"], body_xyz=c["body_xyz"], ds=np.float64(c["ds"]), imu_en=np.int32(imu_en), pose_init=pose_arr(p),
               pose_gt=pose_arr(c["pose_gt"]))
    H, b, m = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    st = sc.get()
    out.update(s_HtH=H, s_Htr=b, s_m=np.int32(m), s_near_cnt=st["near_cnt"], s_near_xyz=st["near_xyz"], s_near_d2=st["near_d2"],
               s_selected=st["selected"], s_normvec=st["normvec"], s_world=st["world"])
    p2 = scenes.perturb_pose(p, 77, dtheta_deg=0.05, dpos=0.01)
    out["pose_2"] = pose_arr(p2)
    H2, b2, m2 = sc.iterate(om, p2.rot_end, p2.pos_end, p2.R_LI, p2.T_LI, imu_en, False)
    st2 = sc.get()
    out.update(r_HtH=H2, r_Htr=b2, r_m=np.int32(m2), r_selected=st2["selected"], r_normvec=st2["normvec"])
    gt = c["pose_gt"]
    cnt, na, nn, flags = sc.map_incremental(om, gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"])
    live = om.flatten()
    live = live[np.lexsort((live[:, 2], live[:, 1], live[:, 0]))]
    out.update(mi_n_add=np.int32(na), mi_n_nod=np.int32(nn), mi_flags=flags.astype(np.int8), mi_live=live)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "m", m, m2, "add", na, nn, "live", len(live))


if __name__ == "__main__":
    one("c1_lo", scenes.make_config("C1", imu_en=False), False)
    one("c1_lio", scenes.make_config("C1", imu_en=True), True)
    one("c2small_lo", scenes.make_config("C2", N=3000, M=40000, open_air_frac=0.02, imu_en=False), False)
    one("c2small_lio", scenes.make_config("C2", N=3000, M=40000, open_air_frac=0.02, imu_en=True), True)
