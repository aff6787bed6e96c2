"""BASELINE config 4 soak: grow the device map from 5M to ~50M points through the incremental insert path
(Add_Points with downsampling, 1M-point batches along a long corridor world), then time search passes of a
260k-point scan against the big map (800 MB of points: no longer L2 resident)."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
from lidar_imu_init_b200 import scenes, capi

target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
ds = 0.15
rng = np.random.default_rng(5)
g = capi.LiInitGpu(ds, max_map_points=target + 2_000_000, max_scan_points=300_000)
# corridor world: floor + two walls + ceiling, 12 m wide, 6 m high, as long as needed; one ds-cell jittered grid
W, H = 12.0, 6.0
per_m = (2 * W + 2 * H) / (ds * ds)            # points per metre of corridor
chunk_len = 1_000_000 / per_m
log = []
x0 = 0.0
first = True
t_all = time.time()
while True:
    n_live = g.map_validnum()
    if n_live >= target:
        break
    L = chunk_len
    sc = scenes.Scene([scenes.Rect(np.array([x0, 0, 0.0]), np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), L, W),
                       scenes.Rect(np.array([x0, 0, H]), np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), L, W),
                       scenes.Rect(np.array([x0, 0, 0.0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0]), L, H),
                       scenes.Rect(np.array([x0, W, 0.0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0]), L, H)], L, W, H)
    pts = scenes.map_points(sc, ds, None, seed=int(x0) + 1)
    t = time.time()
    if first:
        g.map_build(pts)
        first = False
    else:
        g.map_add_points(pts, True)
    dt = time.time() - t
    log.append((g.map_validnum(), len(pts), dt))
    x0 += L
print(f"grown to {g.map_validnum()} points in {time.time()-t_all:.1f}s wall ({len(log)} batches); stats {g.map_stats()}", flush=True)
ins = np.array([l[2] for l in log[1:]])
print(f"Add_Points(1M-batch, downsample) wall ms: median {1e3*np.median(ins):.1f}, last {1e3*ins[-1]:.1f}", flush=True)
# a 260k-point scan in the middle of the corridor, perturbed pose
mid = 0.5 * x0
scene = scenes.Scene([scenes.Rect(np.array([mid - 60, 0, 0.0]), np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), 120, W),
                      scenes.Rect(np.array([mid - 60, 0, H]), np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), 120, W),
                      scenes.Rect(np.array([mid - 60, 0, 0.0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0]), 120, H),
                      scenes.Rect(np.array([mid - 60, W, 0.0]), np.array([1.0, 0, 0]), np.array([0, 0, 1.0]), 120, H)], 120, W, H)
gt = scenes.Pose(scenes.rot_from_rpy(0.01, -0.02, 0.3), np.array([mid, 6.0, 1.5]), np.eye(3), np.zeros(3))
body = scenes.scan_points(scene, gt, 260_000, seed=9, det_range=150.0, sigma=0.01, open_air_frac=0.0)
p = scenes.perturb_pose(gt, 3)
g.scan_upload(body)
for search in (True, False):
    ts = []
    for _ in range(8):
        H_, b_, m, _ = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, search)
        ts.append(g.last_pass_timing()[0])
    print(f"search={int(search)} pass kernel ms median {np.median(ts[2:]):.4f} m={m}", flush=True)
t = time.time(); na, nn = g.map_incremental(gt.rot_end, gt.pos_end, gt.R_LI, gt.T_LI, ds); print(f"map_incremental 260k pts: {1e3*(time.time()-t):.2f} ms wall, add {na} nods {nn}")
print(json.dumps({"map_points": g.map_validnum(), "stats": g.map_stats()}))
g.close()
