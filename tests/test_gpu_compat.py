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
    assert env.map.map_level in ("Normal", "Complex", "Extrem") and len(env.map.obstacles) > 16  # ParkingMapDLP.reset labels the case (:86)
    assert env.map.case_id == 3
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


class _ScriptedAgent(object):
    """stands in for ParkingAgent in the access pattern of evaluation/eval_utils.py:31-84 (policy = masked random action,
    RS plan executed open loop once the env hands one over, parking_agent.py:12-47)"""

    def __init__(self, rng):
        self.rng, self.plan = rng, []

    def reset(self):
        self.plan = []

    def set_planner_path(self, path):
        if self.plan:
            return
        steer_of = {"L": 1.0, "S": 0.0, "R": -1.0}
        for c, l in zip(path.ctypes, path.lengths):
            rem = l / 1.25
            sgn = 1.0 if rem > 0 else -1.0
            while abs(rem) > 1:
                self.plan.append(np.array([steer_of[c], sgn])); rem -= sgn
            if abs(rem) > 1e-3:
                self.plan.append(np.array([steer_of[c], rem]))

    def choose_action(self, obs):
        if self.plan:
            return self.plan.pop(0), None
        ok = np.flatnonzero(obs["action_mask"] > 0.3)
        j = int(self.rng.choice(ok)) if len(ok) else int(self.rng.integers(0, 42))
        return np.array([1.0 - 0.1 * (j % 21), 1.0 if j < 21 else -1.0]), None


def test_eval_loop_access_pattern_on_every_level(compat_env, golden_dir):
    """evaluation/eval_utils.py:31-84, 98-99 line by line against the facade: env.reset(i + 1), vehicle.state.loc, obs['target'],
    action_space.sample(), info['path_to_dest'] / info['status'], map.case_id / map.map_level, and get_map_level on map.start /
    map.dest / map.obstacles — three episodes on each of the four levels of eval_mix_scene.py:82-109, 'dlp' included."""
    from hope_b200 import dlp
    from env.map_level import get_map_level
    CarParking, CarParkingWrapper, Status, VALID_SPEED = compat_env
    raw = CarParking(fps=100, verbose=False, render_mode="rgb_array", use_img_observation=False)
    raw._dlp_cases = dlp.cases_from_fixture(np.load(os.path.join(golden_dir, "dlp_cases.npz")))  # instead of ../data/dlp.data
    env = CarParkingWrapper(raw)
    env.action_space.seed(42)
    np.random.seed(42)
    agent = _ScriptedAgent(np.random.default_rng(0))
    arrived = 0
    for level in ("Extrem", "dlp", "Complex", "Normal"):
        env.set_level(level)
        for i in range(3):
            obs = env.reset(i + 1)
            agent.reset()
            done, step_num, total_reward, path_length = False, 0, 0, 0
            last_xy = (env.vehicle.state.loc.x, env.vehicle.state.loc.y)
            last_obs = obs["target"]
            while not done:
                step_num += 1
                action, _ = agent.choose_action(obs)
                if (last_obs == obs["target"]).all():
                    action = env.action_space.sample()
                last_obs = obs["target"]
                next_obs, reward, done, info = env.step(action)
                total_reward += reward
                obs = next_obs
                path_length += np.linalg.norm(np.array(last_xy) - np.array((env.vehicle.state.loc.x, env.vehicle.state.loc.y)))
                last_xy = (env.vehicle.state.loc.x, env.vehicle.state.loc.y)
                if info["path_to_dest"] is not None:
                    agent.set_planner_path(info["path_to_dest"])
                if done:
                    arrived += int(info["status"] == Status.ARRIVED)
            assert 1 <= step_num <= 201 and np.isfinite(total_reward) and np.isfinite(path_length)
            assert isinstance(info["status"], Status) and info["status"] != Status.CONTINUE
            if level == "dlp":
                assert env.map.case_id == i + 1 and env.map.map_level in ("Normal", "Complex", "Extrem")
            else:
                assert env.map.case_id in (0, 1) and env.map.map_level == level
            assert get_map_level(env.map.start, env.map.dest, env.map.obstacles) in ("Normal", "Complex", "Extrem")  # eval_utils.py:99
            assert len(env.vehicle.trajectory) >= 1 and env.vehicle.box is not None
    assert arrived >= 1  # the open-loop RS hand-off parks at least once in 12 episodes
    env.close()


def test_step_without_action_render_and_wrapper_hooks(compat_env, golden_dir):
    """CarParking.step(None) (car_parking_base.py:255: no motion, t += 1, full observation), render() returning the current observation,
    and caller-supplied action / reward / observation functions (env_wrapper.py:58-66) evaluated on the host."""
    CarParking, CarParkingWrapper, Status, _ = compat_env
    import env.env_wrapper as wrap
    g = dict(np.load(os.path.join(golden_dir, "episodes_Normal.npz")))
    scene = {k: g["scene_" + k][0:1] for k in ("start", "dest", "bounds", "obs", "nverts", "case_id")}
    raw = CarParking(render_mode="rgb_array", verbose=False, use_img_observation=False)
    raw.load_scene(scene)
    obs0 = raw.reset()
    x0 = (raw.vehicle.state.loc.x, raw.vehicle.state.loc.y, raw.vehicle.state.heading)
    obs1, reward_info, status, info = raw.step()          # no action
    assert raw.t == 2.0 and (raw.vehicle.state.loc.x, raw.vehicle.state.loc.y, raw.vehicle.state.heading) == x0
    assert np.array_equal(obs1["lidar"], obs0["lidar"]) and np.array_equal(obs1["action_mask"], obs0["action_mask"])
    assert status == Status.CONTINUE and reward_info["dist_reward"] == 0.0 and reward_info["time_cost"] == -np.tanh(2 / 2000)
    shown = raw.render("rgb_array")
    assert np.array_equal(shown["lidar"], obs1["lidar"]) and shown["img"] is None
    # CarParking.step takes physical units: the same motion as the wrapper's rescaled policy action
    rows = np.where(g["ep"] == 0)[0][:5]
    env_a = CarParkingWrapper(raw)
    raw.load_scene(scene); env_a.reset()
    via_wrapper = [env_a.step(g["action"][i]) for i in rows]
    raw.load_scene(scene); raw.reset()
    for i, (o, r, d, inf) in zip(rows, via_wrapper):
        a = np.clip(g["action"][i], -1, 1) * np.array([0.75, 2.5])
        o2, ri2, st2, inf2 = raw.step(a)
        assert np.array_equal(o2["lidar"], o["lidar"]) and st2 == inf["status"]
    # custom hooks
    calls = {"a": 0, "r": 0, "o": 0}

    def my_action(action, space):
        calls["a"] += 1
        return wrap.action_rescale(-np.asarray(action), space)

    def my_reward(obs, reward_info, status, info):
        calls["r"] += 1
        info["status"] = status
        return obs, 123.0 + reward_info["time_cost"], status, info

    def my_obs(obs):
        calls["o"] += 1
        obs["extra"] = 1
        return obs

    env_b = CarParkingWrapper(raw, action_func=my_action, reward_func=my_reward, observation_func=my_obs)
    raw.load_scene(scene); ob = env_b.reset()
    assert ob["extra"] == 1
    o, r, d, inf = env_b.step(-g["action"][rows[0]])     # negated twice: the recorded first step
    assert calls == {"a": 1, "r": 1, "o": 2} and r == 123.0 + inf["reward_info"]["time_cost"]
    assert np.array_equal(o["lidar"], via_wrapper[0][0]["lidar"])
    assert isinstance(env_b.step(None), dict) and calls["o"] == 3   # env_wrapper.py:74-75
    raw.close()


def test_device_and_host_api_can_be_mixed_without_a_sync(golden_dir):
    """hope_step on the caller's stream followed directly by hope_step_host on the context's own streams (and back): the host
    call orders itself behind the device call.  Same trajectory as stepping through the host API alone."""
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    n = 2048
    sc = generate_scenes(n, "mix", 5)
    a = BatchedParkingEnv(n, scenes=sc, auto_reset=False)
    b = BatchedParkingEnv(n, scenes=sc, auto_reset=False)
    rng = np.random.default_rng(1)
    a.reset(); b.reset_host()
    for k in range(12):
        act = rng.uniform(-1, 1, size=(n, 2))
        if k % 2 == 0:
            a.step(torch.as_tensor(act, device=a.device))        # asynchronous, no synchronize
            out_a = None
        else:
            out_a = a.step_host(act, outputs=("pose", "lidar", "status"))
        out_b = b.step_host(act, outputs=("pose", "lidar", "status"))
        if out_a is not None:
            assert np.array_equal(out_a["pose"], out_b["pose"]) and np.array_equal(out_a["lidar"], out_b["lidar"])
            assert np.array_equal(out_a["status"], out_b["status"])
    a.close(); b.close()
