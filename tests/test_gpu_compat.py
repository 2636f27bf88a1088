"""GPU test of the N = 1 drop-in facade (hope_b200/compat/env): the reference's class surface,
driven exactly like src/train and src/evaluation drive it, replaying a recorded reference episode."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def compat_env():
    path = os.path.join(ROOT, "hope_b200", "compat")
    sys.path.insert(0, path)
    for m in [k for k in sys.modules if k == "env" or k.startswith("env.")]:
        del sys.modules[m]
    from env.car_parking_base import CarParking
    from env.env_wrapper import CarParkingWrapper
    from env.vehicle import Status, VALID_SPEED
    yield CarParking, CarParkingWrapper, Status, VALID_SPEED
    sys.path.remove(path)
    for m in [k for k in sys.modules if k == "env" or k.startswith("env.")]:
        del sys.modules[m]


def test_surface_matches_what_the_scripts_touch(compat_env):
    CarParking, CarParkingWrapper, Status, VALID_SPEED = compat_env
    raw = CarParking(fps=100, verbose=False, render_mode="rgb_array", use_img_observation=False)
    env = CarParkingWrapper(raw)
    assert env.observation_shape == {"action_mask": (42,), "lidar": (120,), "target": (5,)}
    assert env.action_space.shape == (2,) and np.allclose(env.action_space.high, [0.75, 2.5])
    env.action_space.seed(0)
    assert env.action_space.sample().dtype == np.float32
    # train_HOPE_sac.py:164 — RsPlanner step ratio
    assert raw.vehicle.kinetic_model.step_len * raw.vehicle.kinetic_model.n_step * VALID_SPEED[1] == 1.25
    for level in ("Normal", "Complex", "Extrem"):
        obs = env.reset(None, None, level)
        assert obs["img"] is None and obs["lidar"].shape == (120,) and obs["action_mask"].shape == (42,) and obs["target"].shape == (5,)
        assert obs["lidar"].dtype == np.float64
        assert env.map.map_level == level and env.map.case_id in (0, 1) and len(env.map.obstacles) >= 3
        obs, reward, done, info = env.step(np.array([0.3, 1.0]))
        assert isinstance(reward, float) and isinstance(done, bool)
        assert isinstance(info["status"], Status) and list(info["reward_info"]) == ["time_cost", "rs_dist_reward", "dist_reward", "angle_reward", "box_union_reward"]
        assert np.isfinite(env.vehicle.state.loc.x) and np.isfinite(env.vehicle.state.loc.y)  # eval_utils.py:38
    assert env.reset(0, None, "Normal") is not None and env.map.case_id == 0
    assert env.reset(1, None, "Normal") is not None and env.map.case_id == 1
    env.close()


@pytest.mark.parametrize("level", ["Normal", "Complex"])
def test_replays_a_recorded_reference_episode(compat_env, golden_dir, level):
    CarParking, CarParkingWrapper, Status, _ = compat_env
    g = dict(np.load(os.path.join(golden_dir, f"episodes_follow_{level}.npz")))
    env = CarParkingWrapper(CarParking(render_mode="rgb_array", verbose=False, use_img_observation=False))
    n_found = 0
    for e in range(3):
        env.load_scene({k: g["scene_" + k][e:e + 1] for k in ("start", "dest", "bounds", "obs", "nverts", "case_id")})
        obs = env.reset()
        assert np.abs(obs["lidar"] - g["scene_reset_lidar"][e]).max() < 1e-9
        assert np.array_equal(obs["action_mask"], g["scene_reset_mask"][e])
        for i in np.where(g["ep"] == e)[0]:
            obs, reward, done, info = env.step(g["action"][i])
            assert np.abs(obs["lidar"] - g["lidar"][i]).max() < 1e-9
            assert np.array_equal(obs["action_mask"], g["mask"][i])
            assert np.abs(obs["target"] - g["target"][i]).max() < 1e-9
            assert abs(reward - g["reward"][i]) < 1e-9 and done == bool(g["done"][i])
            assert info["status"].value == g["status"][i]
            st = env.vehicle.state
            assert np.abs(np.array([st.loc.x, st.loc.y, st.heading]) - g["pose"][i]).max() < 1e-9
            p = info["path_to_dest"]
            assert (p is not None) == bool(g["rs_found"][i])
            if p is not None:
                n_found += 1
                want = [{0: "S", 1: "L", 2: "R"}[c] for c in g["rs_types"][i][:g["rs_nseg"][i]]]
                if p.ctypes == want:
                    assert np.abs(np.array(p.lengths) - g["rs_lengths"][i][:len(want)]).max() < 1e-9
    assert n_found > 0
    env.close()


def test_facade_switches_to_dlp_scenes(compat_env, golden_dir):
    """eval_mix_scene.py drives `env.reset(case_id, None, 'dlp')`; the facade reads the cases itself."""
    from hope_b200 import dlp
    CarParking, CarParkingWrapper, Status, _ = compat_env
    raw = CarParking(render_mode="rgb_array", verbose=False, use_img_observation=False)
    raw._dlp_cases = dlp.cases_from_fixture(np.load(os.path.join(golden_dir, "dlp_cases.npz")))  # instead of ../data/dlp.data
    env = CarParkingWrapper(raw)
    obs = env.reset(3, None, "dlp")
    assert env.map.map_level == "dlp" and len(env.map.obstacles) > 16
    assert obs["lidar"].shape == (120,) and (obs["lidar"] < 9.9).any()
    obs, reward, done, info = env.step(np.array([0.0, 0.5]))
    assert np.isfinite(obs["lidar"]).all() and isinstance(info["status"], Status)
    obs = env.reset(None, None, "Normal")   # and back to the 16-ring backend
    assert len(env.map.obstacles) <= 16
    env.close()


def test_image_modality_through_the_facade(compat_env, golden_dir):
    """USE_IMG defaults to True in the reference (configs.py:100): the facade returns the float64 (3, 64, 64)
    image of env_wrapper.py:52-55; replay of an episode recorded from the unmodified reference."""
    CarParking, CarParkingWrapper, Status, _ = compat_env
    g = dict(np.load(os.path.join(golden_dir, "images_Normal.npz")))
    raw = CarParking(fps=100, verbose=False, render_mode="rgb_array")  # image on by default
    env = CarParkingWrapper(raw)
    assert env.observation_shape == {"action_mask": (42,), "img": (3, 64, 64), "lidar": (120,), "target": (5,)}
    assert raw.observation_space["img"].shape == (64, 64, 3)
    ep = 1
    rows = np.where(g["ep"] == ep)[0]
    raw.load_scene(dict(start=g["scene_start"][ep:ep + 1], dest=g["scene_dest"][ep:ep + 1], bounds=g["scene_bounds"][ep:ep + 1],
                        obs=g["scene_obs"][ep:ep + 1], nverts=g["scene_nverts"][ep:ep + 1]))
    obs = env.reset()
    assert obs["img"].shape == (3, 64, 64) and obs["img"].dtype == np.float64
    assert np.array_equal(obs["img"], g["scene_reset_img"][ep] / 255.0)
    for k in rows:
        obs, reward, done, info = env.step(g["action"][k])
        assert np.array_equal(obs["img"], g["img"][k] / 255.0), f"step {k}"
        assert 0.0 <= obs["img"].min() and obs["img"].max() <= 1.0
        if done:
            break
    env.close()
