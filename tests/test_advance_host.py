"""k_advance's per-env code replayed on the CPU against traces of the unmodified reference.

hope_b200/csrc/advance_body.inc (what each thread of k_advance runs: action rescale, the ten kinematic substeps with
arrival / collision / retreat, status priority, reward and its shaping, target representation) and advance.cuh's
warp-pooled ring-vs-ring collision test are compiled with g++ by tests/advance_host_harness.cpp on the 32-fiber warp
emulation of tests/warp_emu.h.  All recorded episodes of tests/golden/episodes*_*.npz (oracle/make_golden.py, from
`CarParkingWrapper.step`, env_wrapper.py:73-81) run side by side as the lanes of two emulated warps, free-running from the
reset with the recorded float64 actions, like tests/test_gpu_parity.py::test_golden_episodes_through_cuda does on the
GPU: collision / retreat counts, substeps and status must be identical, pose within 1e-9 (spec 1e-5).
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LEVELS = ("Normal", "Complex", "Extrem")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("advance") / "advance_host.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(HERE, "host_stubs"),
                           "-o", out, os.path.join(HERE, "advance_host_harness.cpp")], env=env)
    lib = C.CDLL(out)
    lib.adv_create.argtypes = [C.c_int, C.c_void_p]
    lib.adv_set_scene.argtypes = [C.c_int] + [C.c_void_p] * 5
    lib.adv_launch.argtypes = [C.c_void_p, C.c_int]
    lib.adv_read.argtypes = [C.c_void_p] * 9
    return lib


def read(lib, n):
    o = dict(pose=np.zeros((n, 3)), status=np.zeros(n, dtype=np.int32), reward=np.zeros(n), reward_info=np.zeros((n, 5)),
             target=np.zeros((n, 5)), substeps=np.zeros(n, dtype=np.uint8), retreated=np.zeros(n, dtype=np.uint8),
             done=np.zeros(n, dtype=np.uint8))
    fb = C.c_ulonglong(0)
    lib.adv_read(*[o[k].ctypes.data for k in ("pose", "status", "reward", "reward_info", "target", "substeps", "retreated", "done")], C.byref(fb))
    o["exact_fallbacks"] = fb.value
    return o


def test_recorded_episodes_free_running(harness, golden_dir):
    from hope_b200 import capi
    par = capi.Params()
    capi.check(capi.load_library().hope_default_params(C.byref(par)))  # host-only call
    files = [dict(np.load(os.path.join(golden_dir, f"{stem}_{lv}.npz"))) for stem in ("episodes", "episodes_follow") for lv in LEVELS]
    eps = [(g, e) for g in files for e in range(len(g["scene_start"]))]   # one lane per recorded episode
    n = len(eps)
    assert 33 <= n <= 64  # two emulated warps, the second with shadow lanes at the tail
    assert harness.adv_create(n, C.addressof(par)) == 0
    for i, (g, e) in enumerate(eps):
        f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        s, d, b, o = f8(g["scene_start"][e]), f8(g["scene_dest"][e]), f8(g["scene_bounds"][e]), f8(g["scene_obs"][e])
        nv = np.ascontiguousarray(g["scene_nverts"][e], dtype=np.int32)
        assert harness.adv_set_scene(i, s.ctypes.data, d.ctypes.data, b.ctypes.data, o.ctypes.data, nv.ctypes.data) == 0
    assert harness.adv_launch(None, 1) == 0
    out = read(harness, n)
    reset_target = np.array([g["scene_reset_target"][e] for g, e in eps])
    assert np.abs(out["target"] - reset_target).max() < 1e-9
    assert np.abs(out["pose"] - np.array([g["scene_start"][e] for g, e in eps])).max() == 0.0
    idx = [np.where(g["ep"] == e)[0] for g, e in eps]
    worst = dict(pose=0.0, reward=0.0, reward_info=0.0, target=0.0)
    compared, statuses, retreats = 0, set(), 0
    for k in range(max(len(i) for i in idx)):
        live = np.array([k < len(i) for i in idx])
        act = np.zeros((n, 2))
        for lane, ((g, e), rows) in enumerate(zip(eps, idx)):
            if k < len(rows):
                act[lane] = g["action"][rows[k]]
        assert harness.adv_launch(act.ctypes.data, 0) == 0, "warp convergence error in the emulation"
        out = read(harness, n)
        for lane, ((g, e), rows) in enumerate(zip(eps, idx)):
            if k >= len(rows):
                continue
            r = rows[k]
            assert out["status"][lane] == g["status"][r], (lane, k)
            assert out["substeps"][lane] == g["substeps"][r] and out["retreated"][lane] == g["retreated"][r], (lane, k)
            assert bool(out["done"][lane]) == bool(g["done"][r])
            worst["pose"] = max(worst["pose"], float(np.abs(out["pose"][lane] - g["pose"][r]).max()))
            worst["reward"] = max(worst["reward"], abs(float(out["reward"][lane] - g["reward"][r])))
            worst["reward_info"] = max(worst["reward_info"], float(np.abs(out["reward_info"][lane] - g["reward_info"][r]).max()))
            worst["target"] = max(worst["target"], float(np.abs(out["target"][lane] - g["target"][r]).max()))
            statuses.add(int(g["status"][r]))
            retreats += int(g["retreated"][r])
            compared += 1
        assert live.any()
    print(f"\nk_advance body on the CPU: {compared} env-steps of {n} recorded episodes, worst |diff| {worst}, "
          f"exact-orientation fallbacks {out['exact_fallbacks']}")
    # CONTINUE, ARRIVED, OUTBOUND, OUTTIME all occur (a collision ends in a retreat, not in a status: ENV_COLLIDE is False)
    assert compared >= 5000 and {1, 2, 4, 5}.issubset(statuses) and retreats >= 100, (compared, statuses, retreats)
    assert worst["pose"] < 1e-9 and worst["reward"] < 1e-9 and worst["reward_info"] < 1e-9 and worst["target"] < 1e-9, worst
