"""k_observe's per-env warp code replayed on the CPU against lidar / action-mask traces of the unmodified reference.

hope_b200/csrc/observe_body.inc (the statements each warp of k_observe runs: ego-frame edge staging, the 120-ray cast
with its exact culls and shared-reciprocal division, the screened action-mask sweep, the 5-tap post-process) is
compiled with g++ by tests/observe_host_harness.cpp on the 32-fiber warp emulation of tests/warp_emu.h and fed the
poses of tests/golden/episodes_*.npz (recorded by oracle/make_golden.py from `CarParkingWrapper.step`,
env_wrapper.py:73-81).  The action mask (action_mask.py:166-196) must come out identical, the lidar
(lidar_simulator.py:31-135) within 3e-11 (the heading's cos / sin come from the host libm here, numpy's there).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("observe") / "observe_host.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(HERE, "host_stubs"),
                           "-o", out, os.path.join(HERE, "observe_host_harness.cpp")], env=env)
    lib = C.CDLL(out)
    lib.observe_host.restype = C.c_int
    lib.observe_host.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11 + [C.c_double] + [C.c_void_p] * 3
    return lib


@pytest.fixture(scope="module")
def device_tables():
    """The arrays hope_upload_tables puts on the device: the host tables plus what k_table_reduce / k_table_group derive."""
    from hope_b200 import tables
    tb = tables.host_tables()
    ds = tb["dist_star"].reshape(1200, 42, 10)
    pmaxk = np.ascontiguousarray(np.maximum.accumulate(ds, axis=2).transpose(0, 2, 1))  # [rho][k][j]
    pmax = np.ascontiguousarray(pmaxk[:, 9, :].max(axis=1))
    gpmax = np.ascontiguousarray(pmax.reshape(120, 10).max(axis=1))
    return dict(tb, pmaxk=pmaxk, pmax=pmax, gpmax=gpmax)


def observe(lib, tb, pose, obs, nverts):
    nv = np.ascontiguousarray(nverts, dtype=np.uint8)
    nobs = int((nv > 0).sum())
    assert (nv[:nobs] > 0).all()
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    lidar, mask, steps = np.zeros(120), np.zeros(42), np.zeros(42, dtype=np.uint8)
    rc = lib.observe_host(pose.ctypes.data, nobs, obs.ctypes.data, nv.ctypes.data,
                          *[tb[k].ctypes.data for k in ("ray_a", "ray_b", "lidar_base", "mask_base", "w_lo", "w_hi", "pmaxk", "pmax", "gpmax")],
                          10.0, lidar.ctypes.data, mask.ctypes.data, steps.ctypes.data)
    assert rc == 0, "warp convergence error in the emulation"
    return lidar, mask, steps


@pytest.mark.parametrize("stem", ["episodes", "episodes_follow"])
@pytest.mark.parametrize("level", ["Normal", "Complex", "Extrem"])
def test_lidar_and_mask_equal_the_recorded_reference(harness, device_tables, golden_dir, level, stem):
    g = dict(np.load(os.path.join(golden_dir, f"{stem}_{level}.npz")))  # arrays in memory: an NpzFile decompresses on every access
    n = len(g["ep"])
    worst, mask_diff = 0.0, 0
    for ep in range(len(g["scene_start"])):  # observation of reset (car_parking_base.py:127-138)
        lidar, mask, _ = observe(harness, device_tables, g["scene_start"][ep], g["scene_obs"][ep], g["scene_nverts"][ep])
        worst = max(worst, float(np.abs(lidar - g["scene_reset_lidar"][ep]).max()))
        mask_diff += int(not np.array_equal(mask, g["scene_reset_mask"][ep]))
    for k in range(n):
        ep = int(g["ep"][k])
        lidar, mask, steps = observe(harness, device_tables, g["pose"][k], g["scene_obs"][ep], g["scene_nverts"][ep])
        worst = max(worst, float(np.abs(lidar - g["lidar"][k]).max()))
        mask_diff += int(not np.array_equal(mask, g["mask"][k]))
        assert np.array_equal(mask, np.full(42, 0.01)) if steps.sum() == 0 else np.array_equal(mask, steps / 10)
    assert mask_diff == 0, f"{mask_diff} of {n} action masks differ from the reference's"
    assert worst < 3e-11, worst


def test_adversarial_scenes_against_the_oracle(harness, device_tables):
    """The C oracle is pinned on the reference's traces; here it and the product's k_observe code (both on the host libm)
    see scenes the generators never produce.  Mask step counts must be identical and the lidar bit-identical."""
    from oracle import parking_oracle as po
    from oracle.adversarial import adversarial_scenes
    rng = np.random.default_rng(77)  # other scenes than the recorded ones (seed 2024)
    n = 500
    start, obs, nverts = adversarial_scenes(rng, n)
    dest = start + np.array([5.0, 5.0, 0.0])
    bounds = np.tile(np.array([-100.0, 100.0, -100.0, 100.0]), (n, 1))
    ref = po.OracleEnv(start, dest, bounds, obs, nverts).reset_step()
    lidar_diff, mask_diff, hits = 0, 0, 0
    for i in range(n):
        lidar, mask, steps = observe(harness, device_tables, start[i], obs[i], nverts[i])
        lidar_diff += int(not np.array_equal(lidar, ref["lidar"][i]))
        mask_diff += int(not np.array_equal(steps, ref["mask_steps"][i].astype(np.uint8)))
        hits += int((lidar < lidar.max() - 1e-9).sum())
    assert hits > 20 * n  # the scenes are in view
    assert mask_diff == 0, f"{mask_diff} of {n} masks differ"
    assert lidar_diff == 0, f"{lidar_diff} of {n} lidar vectors are not bit-identical"


def test_adversarial_scenes_against_the_recorded_reference(harness, device_tables, golden_dir):
    """The same kind of boundary-case scenes, observed by the unmodified reference's LidarSimlator / ActionMask
    (oracle/make_adversarial_golden.py).  The product's k_observe code and the C oracle must both reproduce them."""
    from oracle import parking_oracle as po
    g = dict(np.load(os.path.join(golden_dir, "adversarial_observe.npz")))
    n = len(g["start"])
    ref = po.OracleEnv(g["start"], g["start"] + np.array([5.0, 5.0, 0.0]), np.tile(np.array([-100.0, 100.0, -100.0, 100.0]), (n, 1)),
                       g["obs"], g["nverts"]).reset_step()
    worst_p, worst_o, mask_p, mask_o, exact = 0.0, 0.0, 0, 0, 0
    for i in range(n):
        lidar, mask, _ = observe(harness, device_tables, g["start"][i], g["obs"][i], g["nverts"][i])
        worst_p = max(worst_p, float(np.abs(lidar - g["lidar"][i]).max()))
        worst_o = max(worst_o, float(np.abs(ref["lidar"][i] - g["lidar"][i]).max()))
        mask_p += int(not np.array_equal(mask, g["mask"][i]))
        mask_o += int(not np.array_equal(ref["mask"][i], g["mask"][i]))
        exact += int(np.array_equal(lidar, g["lidar"][i]))
    print(f"\nadversarial scenes vs the reference: product lidar worst |diff| {worst_p:.3g} ({exact} of {n} bit-identical), "
          f"oracle {worst_o:.3g}; masks differing: product {mask_p}, oracle {mask_o}")
    assert mask_p == 0 and mask_o == 0
    assert worst_p < 3e-11 and worst_o < 3e-11
