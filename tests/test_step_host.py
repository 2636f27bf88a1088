"""The whole env step as the kernels run it, chained on the CPU and stepped in lock step with the C oracle.

tests/step_host_harness.cpp strings the product's device source together for test purposes — k_advance's body (32 envs per
emulated warp), k_observe's body (a warp per env), enumerate_env, plan_word, k_rs_check's warp code (the shipped pooled line-pair
test, or with HOPE_CHK_POOLED=0 the per-lane edge loop) and k_rs_select's body — and keeps
the state arrays between steps.  This is the CPU counterpart of tests/test_gpu_parity.py::test_full_step_mixed_levels_vs_oracle:
generated scenes of all three levels, uniform random actions, every output of every step against the oracle (which is
pinned on the reference's traces by tests/test_oracle_golden.py).  Both run on the host libm, so floats agree much closer
than on the GPU.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def build(tmp, pooled=1, max_obs=16):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp / f"step_host_{pooled}_{max_obs}.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(HERE, "host_stubs"),
                           f"-DHOPE_CHK_POOLED={pooled}", f"-DHOPE_MAX_OBS={max_obs}", "-o", out, os.path.join(HERE, "step_host_harness.cpp")], env=env)
    lib = C.CDLL(out)
    lib.step_create.argtypes = [C.c_int] + [C.c_void_p] * 10
    lib.step_set_scene.argtypes = [C.c_int] + [C.c_void_p] * 5
    lib.step_launch.argtypes = [C.c_void_p, C.c_int]
    lib.step_read.argtypes = [C.c_void_p] * 17
    return lib


OUT = [("pose", np.float64, (3,)), ("status", np.int32, ()), ("reward", np.float64, ()), ("reward_info", np.float64, (5,)),
       ("target", np.float64, (5,)), ("substeps", np.uint8, ()), ("retreated", np.uint8, ()), ("lidar", np.float64, (120,)),
       ("mask", np.float64, (42,)), ("mask_steps", np.uint8, (42,)), ("rs_found", np.uint8, ()), ("rs_nseg", np.uint8, ()),
       ("rs_types", np.uint8, (5,)), ("rs_lengths", np.float64, (5,)), ("rs_L", np.float64, ()), ("rs_ncand", np.uint8, ()),
       ("rs_ntried", np.uint8, ())]


def read(lib, n):
    o = {k: np.zeros((n,) + s, dtype=dt) for k, dt, s in OUT}
    lib.step_read(*[o[k].ctypes.data for k, _, _ in OUT])
    return o


def _run(lib, sc, steps, min_steps, min_words, min_found):
    from hope_b200 import capi, tables
    from oracle import parking_oracle as po
    par = capi.Params()
    capi.check(capi.load_library().hope_default_params(C.byref(par)))
    tb = tables.host_tables()
    ds = tb["dist_star"].reshape(1200, 42, 10)
    pmaxk = np.ascontiguousarray(np.maximum.accumulate(ds, axis=2).transpose(0, 2, 1))
    pmax = np.ascontiguousarray(pmaxk[:, 9, :].max(axis=1))
    gpmax = np.ascontiguousarray(pmax.reshape(120, 10).max(axis=1))
    n = len(sc["start"])
    assert lib.step_create(n, C.addressof(par), *[a.ctypes.data for a in (tb["ray_a"], tb["ray_b"], tb["lidar_base"], tb["mask_base"],
                                                                              tb["w_lo"], tb["w_hi"], pmaxk, pmax, gpmax)]) == 0
    for i in range(n):
        f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        s, d, b, o = f8(sc["start"][i]), f8(sc["dest"][i]), f8(sc["bounds"][i]), f8(sc["obs"][i])
        nv = np.ascontiguousarray(sc["nverts"][i], dtype=np.int32)
        assert lib.step_set_scene(i, s.ctypes.data, d.ctypes.data, b.ctypes.data, o.ctypes.data, nv.ctypes.data) == 0
    orc = po.OracleEnv(sc["start"], sc["dest"], sc["bounds"], sc["obs"], sc["nverts"])
    # Word totals: the product sums |lengths| left to right (CPython 3.8, the reference's README); the oracle defaults to
    # CPython 3.12's compensated sum() because the golden traces were recorded under 3.12.  The two differ in the last ulp
    # of L, which decides the pop order of mirror-image words.  Put the oracle on the 3.8 rule for this comparison.
    olib = po.lib(int(np.asarray(sc["nverts"]).shape[1]))
    olib.orc_set_py_sum(0)
    try:
        _lockstep(lib, orc, n, steps, min_steps, min_words, min_found)
    finally:
        olib.orc_set_py_sum(1)  # other tests replay the 3.12 recording


MODES = [1, 0]
MODE_IDS = ["shipped_pooled_line_pairs", "per_lane_edge_loop"]


@pytest.mark.parametrize("pooled", MODES, ids=MODE_IDS)
def test_full_step_lockstep_with_the_oracle(tmp_path_factory, pooled):
    from hope_b200.batched_env import generate_scenes
    lib = build(tmp_path_factory.mktemp("step_host"), pooled)
    n, steps, seed = 96, 40, 321
    if os.environ.get("HOPE_STEP_SOAK"):  # longer run by hand: HOPE_STEP_SOAK="envs,steps,seed"
        n, steps, seed = (int(v) for v in os.environ["HOPE_STEP_SOAK"].split(","))
    _run(lib, generate_scenes(n, "mix", seed), steps, 2500, 3000, 50)


@pytest.mark.parametrize("pooled", MODES, ids=MODE_IDS)
def test_full_step_lockstep_on_dragon_lake_scenes(tmp_path_factory, golden_dir, pooled):
    """The same on the 128-ring build (-DHOPE_MAX_OBS=128) with Dragon Lake Parking scenes (31-119 obstacle rings per scene,
    tests/golden/dlp_cases.npz): long obstacle loops, several queue groups per round in the pooled variant."""
    from hope_b200 import dlp
    lib = build(tmp_path_factory.mktemp("step_host_dlp"), pooled, max_obs=128)
    cases = dlp.cases_from_fixture(np.load(os.path.join(golden_dir, "dlp_cases.npz")))
    sc = dlp.prepare_scenes(cases, np.arange(48) % 16, seed=11)
    _run(lib, sc, 20, 500, 50, 0)  # the recorded starts are mostly farther than 10 m from the slot: few searches


def _lockstep(lib, orc, n, steps, min_steps, min_words, min_found):
    assert lib.step_launch(None, 1) >= 0
    ref = orc.reset_step()
    out = read(lib, n)
    assert np.array_equal(out["lidar"], ref["lidar"]) and np.array_equal(out["mask_steps"], ref["mask_steps"].astype(np.uint8))
    rng = np.random.default_rng(99)
    live = np.ones(n, dtype=bool)
    compared = words = found = other_word = found_flips = 0
    worst = {}
    for _ in range(steps):
        act = np.ascontiguousarray(rng.uniform(-1.0, 1.0, size=(n, 2)))
        n_items = lib.step_launch(act.ctypes.data, 0)
        assert n_items >= 0, "warp convergence error in the emulation"
        ref = orc.step(act)
        out = read(lib, n)
        for key in ("status", "substeps", "retreated", "rs_ncand"):
            assert np.array_equal(out[key][live], ref[key][live].astype(out[key].dtype)), key
        # a word that just grazes an obstacle flips with the last bits of the pose, in the reference itself
        # (tests/test_rs_search_host.py::test_grazing_cases_follow_the_reference_at_both_poses): counted, not required
        found_flips += int((out["rs_found"][live] != ref["rs_found"][live]).sum())
        assert np.array_equal(out["mask_steps"][live], ref["mask_steps"][live].astype(np.uint8))
        # The free-running poses differ in the last bits (closed-form position sum), which is enough to
        # flip the pop order of mirror-image words whose lengths are equal in exact arithmetic: the number of words tried
        # and, when both are clean, the word chosen may then differ (the reference itself derives that order from rounding
        # noise).  From identical poses the search is identical: tests/test_rs_search_host.py.
        both = live & (out["rs_found"] == 1) & (ref["rs_found"] == 1)
        same = both & (out["rs_types"] == ref["rs_types"]).all(axis=1)
        other_word += int(both.sum() - same.sum())
        for key, want, sel in (("pose", orc.pose, live), ("lidar", ref["lidar"], live), ("mask", ref["mask"], live), ("target", ref["target"], live),
                               ("reward", ref["reward"], live), ("reward_info", ref["reward_info"], live), ("rs_lengths", ref["rs_len"], same),
                               ("rs_L", ref["rs_L"], same)):
            d = np.abs(out[key][sel] - want[sel])
            if d.size:
                worst[key] = max(worst.get(key, 0.0), float(d.max()))
        compared += int(live.sum()); words += n_items; found += int(out["rs_found"][live].sum())
        live &= ref["status"] == 1  # the oracle env has no auto-reset: stop comparing finished episodes
    print(f"\nfull step on the CPU: {compared} env-steps, {words} tried words, "
          f"{found} paths found ({other_word} with another of two equal-length words, {found_flips} found/not-found flips), worst |diff| {worst}")
    assert compared >= min_steps and words >= min_words and found >= min_found and other_word <= max(2, found // 20), (compared, words, found, other_word)
    assert found_flips <= max(2, compared // 5000), found_flips
    assert all(v < 1e-9 for v in worst.values()), worst
