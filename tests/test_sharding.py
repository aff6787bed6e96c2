"""N > 1 host logic on CPU: world_size-2 gloo. Each rank evaluates ITS shard of the scan (with the CPU oracle standing
in for the device pass), the 160-double accumulator block is all-reduced, and the sum must equal the single-process
result over the whole scan -- the property the multi-GPU path relies on."""
import os
import socket

import numpy as np
import pytest

from lidar_imu_init_b200 import scenes, sharding


def test_shard_bounds_partition():
    for n in (0, 1, 7, 240000, 2000001):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = sharding.shard_bounds(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
            sizes = [sharding.shard_bounds(n, r, world)[1] - sharding.shard_bounds(n, r, world)[0] for r in range(world)]
            cnt = -(-n // world)
            assert all(s <= cnt for s in sizes) and all(sharding.shard_bounds(n, r, world)[0] == min(r * cnt, n) for r in range(world))   # equal slots
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = scenes.make_config("C2", N=3001, M=30000, open_air_frac=0.02, imu_en=True)
    p = c["pose_init"]
    om = orc.OracleMap(c["ds"], 0)
    om.build(c["map_xyz"])                                   # map replicated on every rank
    lo, hi = sharding.shard_bounds(len(c["body_xyz"]), rank, world)
    sc = orc.OracleScan(c["body_xyz"][lo:hi])                # this rank's shard
    H, b, m = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True, nthreads=1)
    _, meas, _ = sc.get_H()
    acc = torch.zeros(sharding.ACC_DOUBLES, dtype=torch.float64)
    acc[:144] = torch.from_numpy(H.reshape(-1))
    acc[144:156] = torch.from_numpy(b)
    acc[156] = float((meas ** 2).sum())
    acc[157] = m
    sharding.allreduce_accumulators(acc)
    np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0, a1 = np.load(tmp_path / "acc0.npy"), np.load(tmp_path / "acc1.npy")
    assert np.array_equal(a0, a1)                            # every rank holds the same sum
    c = scenes.make_config("C2", N=3001, M=30000, open_air_frac=0.02, imu_en=True)
    p = c["pose_init"]
    om = oracle_mod.OracleMap(c["ds"], 0)
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(c["body_xyz"])
    H, b, m = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, True, True, nthreads=1)
    Hs, bs, rs, ms = sharding.unpack_accumulators(a0)
    assert ms == m
    assert np.allclose(Hs, H, rtol=1e-12, atol=0) and np.allclose(bs, b, rtol=1e-10, atol=1e-12)
