"""CPU tests: the C oracle (oracle/c/parking_oracle.c) against traces recorded from the
UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle.

Bars: integers / booleans / type codes exact; float64 observations bit-exact except where the
recording went through numpy's SIMD tan/tanh (reward time_cost) — 1e-12 there."""
import hashlib
import os

import numpy as np
import pytest

from oracle import parking_oracle as po

LEVELS = ("Normal", "Complex", "Extrem")


def test_mask_tables_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_table.npz"))
    tb = po.mask_tables()
    assert np.array_equal(po.discrete_actions(), g["discrete_actions"])
    assert np.array_equal(po.swept_boxes(), g["vehicle_boxes"])
    assert np.array_equal(tb["mask_base"], g["vehicle_lidar_base"])
    assert np.array_equal(tb["lidar_base"], g["vehicle_boundary"])
    assert tuple(g["dist_star_shape"]) == tb["dist_star"].shape == (1200, 42, 10)
    assert np.array_equal(tb["dist_star"].reshape(-1)[::int(g["dist_star_stride"])], g["dist_star_sample"])
    assert hashlib.sha256(tb["dist_star"].tobytes()).digest() == g["dist_star_sha256"].tobytes()


def test_reeds_shepp_known_answers(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "reeds_shepp.npz")))
    for i in range(len(g["q"])):
        r = po.rs_all_paths(g["q"][i, :3], g["q"][i, 3:], float(g["maxc"]))
        k = int(g["npaths"][i])
        assert r["n"] == k and not r["err"]
        assert np.array_equal(r["nseg"], g["nseg"][i, :k])
        assert np.array_equal(r["types"], g["types"][i, :k])
        assert np.array_equal(r["lengths"], g["lengths"][i, :k])
        assert np.array_equal(r["L"], g["L"][i, :k])
        assert np.array_equal(r["T"], g["T"][i, :k])
        assert np.array_equal(r["head"], g["head"][i, :k])
        assert np.array_equal(r["tail"], g["tail"][i, :k])
        np.testing.assert_allclose(r["csum"], g["csum"][i, :k], rtol=0, atol=1e-7)  # fsum vs plain sum


def _replay(path):
    g = dict(np.load(path))
    ep = g["ep"]
    n_steps = 0
    for e in range(len(g["scene_start"])):
        sl = slice(e, e + 1)
        env = po.OracleEnv(g["scene_start"][sl], g["scene_dest"][sl], g["scene_bounds"][sl], g["scene_obs"][sl],
                           g["scene_nverts"][sl], nthreads=1)
        o = env.reset_step()
        assert np.array_equal(o["lidar"][0], g["scene_reset_lidar"][e])
        assert np.array_equal(o["mask"][0], g["scene_reset_mask"][e])
        np.testing.assert_allclose(o["target"][0], g["scene_reset_target"][e], rtol=0, atol=1e-14)
        for i in np.where(ep == e)[0]:
            o = env.step(g["action"][i:i + 1])
            n_steps += 1
            msg = f"{os.path.basename(path)} episode {e} step {i}"
            assert np.array_equal(env.pose[0], g["pose"][i]), msg
            assert o["status"][0] == g["status"][i], msg
            assert o["substeps"][0] == g["substeps"][i] and o["retreated"][0] == g["retreated"][i], msg
            assert np.array_equal(o["lidar"][0], g["lidar"][i]), msg
            assert np.array_equal(o["mask"][0], g["mask"][i]), msg
            np.testing.assert_allclose(o["target"][0], g["target"][i], rtol=0, atol=1e-14, err_msg=msg)
            np.testing.assert_allclose(o["reward"][0], g["reward"][i], rtol=0, atol=1e-12, err_msg=msg)
            np.testing.assert_allclose(o["reward_info"][0], g["reward_info"][i], rtol=0, atol=1e-12, err_msg=msg)
            assert bool(o["status"][0] != 1) == bool(g["done"][i]), msg
            assert o["rs_found"][0] == g["rs_found"][i] and o["rs_nseg"][0] == g["rs_nseg"][i], msg
            assert o["rs_ncand"][0] == g["rs_ncand"][i] and o["rs_ntried"][0] == g["rs_ntried"][i], msg
            assert o["rs_T_last"][0] == g["rs_T_last"][i], msg
            assert np.array_equal(o["rs_types"][0], g["rs_types"][i]), msg
            assert np.array_equal(o["rs_len"][0], g["rs_lengths"][i]), msg
            assert o["rs_L"][0] == g["rs_L"][i], msg
            assert o["rs_err"][0] == 0
    return n_steps


@pytest.mark.parametrize("level", LEVELS)
def test_random_action_episodes(golden_dir, level):
    """BASELINE cfg 1: float64 U(-1,1)^2 actions, 200 steps per episode."""
    assert _replay(os.path.join(golden_dir, f"episodes_{level}.npz")) > 500


@pytest.mark.parametrize("level", LEVELS)
def test_rs_following_episodes(golden_dir, level):
    """Episodes that execute the RS hand-off: covers ARRIVED and the box-union reward."""
    assert _replay(os.path.join(golden_dir, f"episodes_follow_{level}.npz")) > 1000


def test_env_collide_episodes(golden_dir):
    """ENV_COLLIDE = True (configs.py:79): 75 random-action episodes of all three levels recorded from the unmodified reference
    with the flag on; a collision on the first substep ends the episode with status COLLIDED (car_parking_base.py:264-267, 279-282)."""
    lib = po.lib()
    lib.orc_set_env_collide(1)
    try:
        assert _replay(os.path.join(golden_dir, "episodes_collide.npz")) > 3000
    finally:
        lib.orc_set_env_collide(0)
    assert (np.load(os.path.join(golden_dir, "episodes_collide.npz"))["status"] == 3).sum() >= 50


def test_golden_covers_every_status(golden_dir):
    seen = set()
    for level in LEVELS:
        for stem in ("episodes", "episodes_follow"):
            seen |= set(np.load(os.path.join(golden_dir, f"{stem}_{level}.npz"))["status"].tolist())
    assert {1, 2, 4, 5} <= seen  # COLLIDED (3) is unreachable with ENV_COLLIDE=False (configs.py:79) ...
    seen |= set(np.load(os.path.join(golden_dir, "episodes_collide.npz"))["status"].tolist())
    assert {1, 2, 3, 4, 5} <= seen  # ... and comes from the ENV_COLLIDE=True recording


def test_geometry_predicates_agree_with_python_restatement():
    """C predicates vs oracle/geom.py (exact rational orientation) incl. degenerate inputs."""
    import ctypes as C
    from oracle import geom
    lib = po.lib()
    rng = np.random.default_rng(0)
    for _ in range(3000):
        p = rng.integers(-3, 4, size=8).astype(np.float64)
        if rng.random() < 0.5:
            p += rng.normal(0, 1e-15, size=8)
        want = geom.segments_intersect((p[0], p[1]), (p[2], p[3]), (p[4], p[5]), (p[6], p[7]))
        got = lib.orc_seg_hit(p.ctypes.data_as(C.POINTER(C.c_double)))
        assert bool(got) == bool(want), p
    for _ in range(3000):
        a = rng.uniform(-5, 5, size=2); d = rng.uniform(-1, 1, size=2)
        b = a + d; c = a + d * rng.uniform(-2, 2)  # nearly collinear
        assert lib.orc_orient(a[0], a[1], b[0], b[1], c[0], c[1]) == geom.orient(a[0], a[1], b[0], b[1], c[0], c[1])
