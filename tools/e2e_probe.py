"""Where does the end-to-end step go? CPU-side call durations vs device time (developer probe)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from lidar_imu_init_b200 import scenes, capi
c = scenes.make_config("C2")
p = c["pose_init"]; N = len(c["body_xyz"])
g = capi.LiInitGpu(c["ds"], max_map_points=6_000_000, max_scan_points=N + 16)
st = torch.cuda.Stream(); g.set_stream(st.cuda_stream)
g.map_build(c["map_xyz"])
body4 = torch.zeros((N, 4), dtype=torch.float32).pin_memory(); body4[:, :3] = torch.from_numpy(c["body_xyz"])
for _ in range(5):
    g.scan_upload_ptr(body4.data_ptr(), 4, N); g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
tu, ti, tt = [], [], []
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); g.scan_upload_ptr(body4.data_ptr(), 4, N); t1 = time.perf_counter()
    g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True); t2 = time.perf_counter()
    tu.append(t1 - t0); ti.append(t2 - t1); tt.append(t2 - t0)
print(f"upload call {1e6*np.median(tu):.1f} us | iterate call (incl. sync) {1e6*np.median(ti):.1f} us | total {1e6*np.median(tt):.1f} us")
# pure copy time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d = torch.empty((N, 4), dtype=torch.float32, device="cuda")
ts = []
for _ in range(10):
    e0.record(); d.copy_(body4, non_blocking=True); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"plain pinned H2D of {N*16/1e6:.2f} MB: {1e3*np.median(ts):.1f} us -> {N*16/np.median(ts)/1e6:.1f} GB/s")
ti2 = []
for _ in range(30):
    t1 = time.perf_counter(); g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True); ti2.append(time.perf_counter() - t1)
print(f"resident iterate call {1e6*np.median(ti2):.1f} us (kernel {1e3*g.last_pass_timing()[0]:.1f} us)")

# ---- zero-copy attach vs staged upload, end to end (host call -> accumulators back on the host) -----------------
body3 = torch.from_numpy(np.ascontiguousarray(c["body_xyz"])).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(mode, stride, buf):
    ts, ks = [], []
    for i in range(25):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        (g.scan_attach_ptr if mode == "attach" else g.scan_upload_ptr)(buf.data_ptr(), stride, N)
        g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, False, True)
        t1 = time.perf_counter()
        if i >= 5:
            ts.append(t1 - t0); ks.append(g.last_pass_timing()[0])
    return 1e3 * np.median(ts), np.median(ks)
for mode in ("upload", "attach", "upload", "attach"):
    for stride, buf in ((3, body3), (4, body4)):
        t, k = run(mode, stride, buf)
        print(f"e2e {mode:6s} stride {stride}: wall {t:.4f} ms | device pass {k:.4f} ms")
