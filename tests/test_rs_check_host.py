"""k_rs_check's warp-level trajectory check replayed on the CPU against the unmodified reference's verdicts.

hope_b200/csrc/rs_check.cuh (chunk_is_bad / warp_samples_hit: the code each warp of k_rs_check runs per tried word) is
compiled with g++ by tests/rs_check_host_harness.cpp on top of a 32-fiber warp emulation (tests/warp_emu.h, which also
flags lanes that wait at different collectives or leave early, like synccheck).  The fixture tests/golden/traj_valid.npz
holds `CarParking.is_traj_valid` verdicts (car_parking_base.py:452-534) recorded from the unmodified reference by
oracle/make_traj_valid_golden.py: scene, start pose, which word of calc_all_paths, number of samples, verdict.
The shipped pooled test (rs_check_pooled.cuh), both vote placements of the per-lane edge loop (HOPE_CHK_EDGE_EXIT = 1 and 0)
and the trailing-zero path are covered.
"""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
MAXC = math.tan(0.75) / 2.8   # car_parking_base.py:422
RS_STEP = 0.1                 # :424


@pytest.fixture(scope="module", params=[(1, 1), (1, 0), (0, 0)], ids=["pooled", "edge_exit", "obstacle_exit"])
def harness(request, tmp_path_factory):
    """params: (HOPE_CHK_EDGE_EXIT, HOPE_CHK_POOLED).  pooled is the shipped build (rs_check_pooled.cuh: line-pair tests of a round
    pooled over the warp); edge_exit / obstacle_exit are the per-lane edge loops of rs_check.cuh with either vote placement."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    edge_exit, pooled = request.param
    out = str(tmp_path_factory.mktemp("rs_check") / f"rs_check_{edge_exit}_{pooled}.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(HERE, "host_stubs"),
                           f"-DHOPE_CHK_EDGE_EXIT={edge_exit}", f"-DHOPE_CHK_POOLED={pooled}", "-o", out, os.path.join(HERE, "rs_check_host_harness.cpp")], env=env)
    lib = C.CDLL(out)
    lib.rs_check_host.restype = C.c_int
    lib.rs_check_host.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]
    return lib


@pytest.fixture(scope="module")
def golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "traj_valid.npz")))  # arrays in memory: an NpzFile decompresses on every access


@pytest.fixture(scope="module")
def box():
    from hope_b200 import capi
    p = capi.Params()
    capi.check(capi.load_library().hope_default_params(C.byref(p)))  # host-only call
    return np.array(list(p.box_x)), np.array(list(p.box_y))


def run_call(lib, g, box, i, zero_tail=0):
    sc = int(g["call_scene"][i])
    q = np.ascontiguousarray(np.concatenate([g["call_pose"][i], g["scene_dest"][sc]]), dtype=np.float64)
    nv = np.ascontiguousarray(g["scene_nverts"][sc], dtype=np.uint8)
    nobs = int((nv > 0).sum())
    assert (nv[:nobs] > 0).all()  # rings are compact at the front, as hope_set_scene_pool requires
    obs = np.ascontiguousarray(g["scene_obs"][sc], dtype=np.float64)
    bounds = np.ascontiguousarray(g["scene_bounds"][sc], dtype=np.float64)
    ns, nc = C.c_int(0), C.c_ulonglong(0)
    rc = lib.rs_check_host(q.ctypes.data, MAXC, RS_STEP, int(g["call_word"][i]), bounds.ctypes.data, nobs, obs.ctypes.data, nv.ctypes.data,
                           box[0].ctypes.data, box[1].ctypes.data, zero_tail, C.byref(ns), C.byref(nc))
    return rc, ns.value, nc.value


def test_verdicts_equal_the_reference(harness, golden, box):
    g = golden
    n = len(g["call_valid"])
    assert n >= 1000 and int(g["call_valid"].sum()) >= 100 and int((g["call_valid"] == 0).sum()) >= 500
    wrong, t_wrong = [], 0
    for i in range(n):
        rc, ns, _ = run_call(harness, g, box, i)
        assert rc in (0, 1), f"call {i}: harness error {rc} (-2 = warp convergence error, -3 = lanes disagree)"
        if rc != (0 if g["call_valid"][i] else 1):
            wrong.append(i)
        if ns >= 0 and ns != int(g["call_T"][i]):
            t_wrong += 1
    # libm (glibc here, CUDA's on the GPU) may differ from numpy's in the last bit of a sin/cos; a verdict can only flip when a
    # sample lands within that of an edge.  None does in this fixture.
    assert not wrong, f"{len(wrong)} of {n} verdicts differ from is_traj_valid: calls {wrong[:10]}"
    assert t_wrong == 0, f"{t_wrong} words have a different number of samples than the reference's trajectory"


def test_trailing_zero_path_gives_the_same_verdicts(harness, golden, box):
    """With end_lx forced to 0.0 chunk_is_bad takes the trailing-zero path (reeds_shepp.py:501-505: samples whose local
    x is exactly 0.0 are dropped from the END of the trajectory): no early exits, hits on x == 0.0 samples only count if
    a later sample has x != 0.0.  The recorded words all have later samples with x != 0.0, so the verdict must not
    change; this is the only test that runs the non-early branch of warp_samples_hit."""
    g = golden
    idx = np.concatenate([np.flatnonzero(g["call_valid"] == 1)[:60], np.flatnonzero(g["call_valid"] == 0)[:240]])
    for i in idx:
        a = run_call(harness, g, box, int(i), 0)
        b = run_call(harness, g, box, int(i), 1)
        assert a[0] in (0, 1) and b[0] == a[0], (int(i), a, b)
