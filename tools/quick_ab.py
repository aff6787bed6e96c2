"""Developer A/B on one box (not the contract bench): kernel ms of the search pass for the two spatial indexes
(bricks = lockstep lane groups over whole bricks, cells = cell directory + thread per point), plus the map-side cost of
keeping the directory (build, map_incremental). L2 is not flushed here (the map is L2-resident in both cases)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lidar_imu_init_b200 import scenes, capi
import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=240000)
ap.add_argument("--M", type=int, default=5000000)
# index:minb:rho_cells:search[:sched]   index 1 bricks / 2 cells / 3 hybrid; minb only honoured by libraries built with that A/B switch (git history);
# search 1 shells on cells / 2 growing boxes / 3 enumerate + stream; sched s = one block per 128 points, d = LIINIT_CELLS_SCHED=dynamic
ap.add_argument("--variants", default="1:0:2:0,2:6:2:3,2:6:2:3:d,3:6:2:3,2:6:2:2,2:6:2:1")
a = ap.parse_args()
t = time.time(); c = scenes.make_config("C2", N=a.N, M=a.M); print("gen", round(time.time() - t, 2), flush=True)
ref = None
for v in a.variants.split(","):
    f = v.split(":")
    idx, minb, rho, search = f[:4]
    sched = f[4] if len(f) > 4 else "s"
    os.environ["LIINIT_CELLS_MINB"] = minb
    os.environ["LIINIT_CELLS_SEARCH"] = search
    os.environ["LIINIT_CELLS_SCHED"] = "dynamic" if sched == "d" else "static"
    g = capi.LiInitGpu(c["ds"], max_map_points=int(a.M * 1.2), max_scan_points=a.N + 10, knn_index=int(idx), knn_seed_radius_cells=float(rho))
    tb = []
    for rep in range(2):
        t = time.time(); g.map_build(c["map_xyz"]); tb.append(time.time() - t)
    g.scan_upload(c["body_xyz"])
    line = f"index {idx} minb {minb} rho {rho} search {search} sched {sched}: build {min(tb)*1e3:.1f} ms"
    for pose_name in ("init", "gt"):
        p = c["pose_" + pose_name]
        ks, ps = [], []
        for it in range(8):
            H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
            k, pl = g.last_pass_kernel_times(); ks.append(k); ps.append(pl)
        line += f" | {pose_name}: m={m} knn {np.median(ks[2:]):.4f} (min {min(ks):.4f}) plane {np.median(ps[2:]):.4f}"
        if pose_name == "init":
            if ref is None: ref = (H.copy(), b.copy(), m)
            else: line += f" [same m: {m == ref[2]}, HtH rel diff {np.abs(H-ref[0]).max()/np.abs(ref[0]).max():.1e}]"
    gt = c["pose_gt"]
    mi = []
    for rep in range(2):
        g.icp_iterate(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, False, True)
        t = time.perf_counter(); na, nn = g.map_incremental(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, c["ds"]); mi.append((time.perf_counter() - t) * 1e3)
    line += f" | map_incremental {mi[1]:.3f} ms (first {mi[0]:.3f}) added {na},{nn} valid {g.map_validnum()}"
    print(line, flush=True)
    g.close()
