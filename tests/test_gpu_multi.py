"""N > 1 on real GPUs (skipped on a single-GPU box): two ranks, NCCL, every rank holds a replica of the map and its shard of
the scan; the all-reduced accumulators must equal the single-GPU result of the whole scan."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from lidar_imu_init_b200 import capi, scenes, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    c = scenes.make_config("C2", N=30001, M=200000, open_air_frac=0.02, imu_en=True)
    p = c["pose_init"]
    g = capi.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=40000, device_id=rank)
    stream = torch.cuda.Stream(device=rank)
    g.set_stream(stream.cuda_stream)
    g.map_build(c["map_xyz"])
    lo, hi = sharding.shard_bounds(len(c["body_xyz"]), rank, world)
    g.scan_upload(c["body_xyz"][lo:hi])
    acc = torch.zeros(sharding.ACC_DOUBLES, dtype=torch.float64, device=f"cuda:{rank}")
    with torch.cuda.stream(stream):
        g.icp_iterate_device(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True, acc.data_ptr())
        sharding.allreduce_accumulators(acc)
        stream.synchronize()
    np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc.cpu().numpy())
    g.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_allreduce_equals_single_gpu(tmp_path, gpu_lib):
    import torch.multiprocessing as mp
    from lidar_imu_init_b200 import scenes, sharding
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a0, a1 = np.load(tmp_path / "acc0.npy"), np.load(tmp_path / "acc1.npy")
    assert np.array_equal(a0, a1)
    c = scenes.make_config("C2", N=30001, M=200000, open_air_frac=0.02, imu_en=True)
    p = c["pose_init"]
    g = gpu_lib.LiInitGpu(c["ds"], max_map_points=400000, max_scan_points=40000)
    g.map_build(c["map_xyz"])
    g.scan_upload(c["body_xyz"])
    H, b, m, rs = g.icp_iterate(p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True)
    Hs, bs, rss, ms = sharding.unpack_accumulators(a0)
    assert ms == m and np.allclose(Hs, H, rtol=1e-12) and np.allclose(bs, b, rtol=1e-10, atol=1e-12) and abs(rss - rs) <= 1e-10 * rs
    g.close()
