"""Host-side library (IESKF in information form, state algebra): CPU tests against the oracle's literal restatement
of laserMapping.cpp:1080-1087 and common_lib.h:109-151. No GPU compute is invoked."""
import numpy as np
import pytest

from lidar_imu_init_b200 import scenes


@pytest.fixture(scope="module")
def host_lib(gpu_lib):
    from lidar_imu_init_b200 import _build, host
    _build.build_host()
    host.load()
    return host


def test_state_algebra_matches_oracle(host_lib, oracle_mod):
    rng = np.random.default_rng(11)
    assert np.array_equal(host_lib.state_init(), oracle_mod.state_pack())
    for _ in range(20):
        v = rng.normal(size=3)
        v *= rng.uniform(1e-4, 3.0) / np.linalg.norm(v)
        R = np.zeros(9)
        host_lib.load().liinit_so3_exp(v, R)
        assert np.array_equal(R.reshape(3, 3), oracle_mod.so3_exp(v))
        w = np.zeros(3)
        host_lib.load().liinit_so3_log(R, w)
        assert np.array_equal(w, oracle_mod.so3_log(R.reshape(3, 3)))
    s = host_lib.state_init()
    d = rng.normal(size=24) * 0.01
    s2 = host_lib.boxplus(s, d)
    back = host_lib.boxminus(s2, s)
    assert np.allclose(back, d, atol=1e-12)


@pytest.mark.parametrize("imu_en", [False, True])
def test_ieskf_information_form_matches_literal(host_lib, oracle_mod, imu_en):
    c = scenes.make_config("C2", N=2500, M=30000, open_air_frac=0.0, imu_en=imu_en)
    p = c["pose_init"]
    om = oracle_mod.OracleMap(c["ds"], 0)
    om.build(c["map_xyz"])
    sc = oracle_mod.OracleScan(c["body_xyz"])
    HtH, Htr, m = sc.iterate(om, p.rot_end, p.pos_end, p.R_LI, p.T_LI, imu_en, True)
    st0 = host_lib.state_from_pose(p.rot_end, p.pos_end, p.R_LI, p.T_LI)
    prop = st0.copy()
    prop[9:12] += [0.01, -0.02, 0.005]
    s_lit, sol_lit, KH_lit = sc.ieskf_update(st0, prop)
    s_inf, sol_inf, KH_inf = host_lib.ieskf_update(st0, prop, HtH, Htr)
    assert np.allclose(sol_inf, sol_lit, rtol=1e-5, atol=1e-9)
    assert np.allclose(s_inf[:36], s_lit[:36], rtol=0, atol=1e-8)      # pose part: far inside the 1e-3 bar
    assert np.allclose(KH_inf, KH_lit, rtol=1e-4, atol=1e-5)
