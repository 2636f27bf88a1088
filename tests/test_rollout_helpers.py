"""Rollout-side helpers (scope row f2) against recordings of the UNMODIFIED reference (oracle/make_f2_golden.py ->
tests/golden/f2_helpers.npz): StateNorm, ReplayMemory and ActionMask.choose_action's probabilities on the CPU through the
product's torch code, and the RsPlanner / ParkingAgent hand-off through k_planner on the GPU."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from hope_b200 import learner, rollout  # noqa: E402


@pytest.fixture(scope="module")
def g(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "f2_helpers.npz")))


def _feed(norm, g, batch):
    n = len(g["norm_lidar"])
    for lo in range(0, n, batch):
        norm.update({"lidar": torch.as_tensor(g["norm_lidar"][lo:lo + batch]), "target": torch.as_tensor(g["norm_target"][lo:lo + batch])})


@pytest.mark.parametrize("batch", [1, 7, 128, 1536])
def test_running_norm_equals_statenorm(g, batch):
    """state_norm.py:25-46 fed one observation at a time vs the batched Welford merge, any batch size: same mean / std / n, and the
    same normalised probe observation."""
    norm = rollout.RunningNorm({"lidar": (120,), "target": (5,)}, "cpu")
    _feed(norm, g, batch)
    assert norm.n == int(g["norm_n"])
    for k in ("lidar", "target"):
        std = torch.sqrt(norm.m2[k] / norm.n).numpy()
        np.testing.assert_allclose(norm.mean[k].numpy(), g[f"norm_mean_{k}"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(std, g[f"norm_std_{k}"], rtol=1e-10, atol=1e-12)
    out = norm({"lidar": torch.as_tensor(g["norm_probe_lidar"])[None], "target": torch.as_tensor(g["norm_probe_target"])[None]})
    np.testing.assert_allclose(out["lidar"][0].numpy(), g["norm_probe_out_lidar"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(out["target"][0].numpy(), g["norm_probe_out_target"], rtol=1e-9, atol=1e-10)


def test_running_norm_follows_statenorm_step_by_step(g):
    """with one env the batched code IS the reference's recurrence, quirks included: the first observation normalises to 0
    (state_norm.py:28-33 sets mean = std = the observation), the second to +-1 / (1 + 2e-8 / |delta|)"""
    norm = rollout.RunningNorm({"lidar": (120,), "target": (5,)}, "cpu")
    first = {"lidar": torch.as_tensor(g["norm_lidar"][0:1]), "target": torch.as_tensor(g["norm_target"][0:1])}
    norm.update(first)
    assert (norm(first)["lidar"] == 0).all()
    second = {"lidar": torch.as_tensor(g["norm_lidar"][1:2]), "target": torch.as_tensor(g["norm_target"][1:2])}
    norm.update(second)
    out = norm(second)
    np.testing.assert_allclose(out["lidar"][0].numpy(), g["norm_second_out_lidar"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["target"][0].numpy(), g["norm_second_out_target"], rtol=1e-12, atol=1e-12)


def test_device_replay_keeps_what_replaymemory_keeps(g):
    """replay_memory.py:6-50 (deque of capacity 96 after 250 pushes) vs the ring buffer: the same transitions survive, with the same
    reward / action / done, and the stored next observation is the reference's state[idx + 1] wherever that exists."""
    cap, pushes = int(g["replay_capacity"]), int(g["replay_pushes"])
    rep = learner.DeviceReplay(cap, "cpu", keys=(("lidar", 2),))
    for i in range(pushes):
        obs = {"lidar": torch.tensor([[float(i), 0.5 * i]])}
        nxt = {"lidar": torch.tensor([[float(i + 1), 0.5 * (i + 1)]])}
        rep.push(obs, torch.tensor([[0.1 * i, -0.1 * i]]), torch.tensor([float(i) * 0.01]), torch.tensor([float(g["replay_done"][i])]), nxt)
    assert rep.size == int(g["replay_len"]) == cap
    held = sorted(int(v) for v in rep.obs["lidar"][:, 0].tolist())
    assert held == sorted(g["replay_ids"].tolist())                    # the newest `capacity` transitions
    for pos, tid in enumerate(g["replay_ids"]):
        slot = int(tid) % cap
        assert int(rep.obs["lidar"][slot, 0]) == tid
        assert abs(float(rep.reward[slot]) - g["replay_reward"][pos]) < 1e-6
        np.testing.assert_allclose(rep.action[slot].numpy(), g["replay_action"][pos], rtol=1e-6)
        assert bool(rep.done[slot]) == bool(g["replay_done"][tid])
        if g["replay_has_next"][pos]:  # the reference derives it from the neighbouring slot; None after a done or at the newest entry
            assert int(rep.nxt["lidar"][slot, 0]) == g["replay_next_id"][pos] == tid + 1
        else:
            assert bool(rep.done[slot]) or tid == pushes - 1
    gen = torch.Generator(); gen.manual_seed(0)
    o, a, r, d, n = rep.sample(4096, gen)
    assert set(int(v) for v in o["lidar"][:, 0].tolist()) == set(held)  # uniform over everything held (:35-37)


def test_masked_action_probabilities_equal_choose_action(g):
    """action_mask.py:199-227: the probabilities np.random.choice is called with, for 200 (mean, std, mask) triples"""
    acts = rollout.possible_actions("cpu")
    p = rollout.masked_action_probs(torch.as_tensor(g["choose_mean"]), torch.as_tensor(g["choose_std"]), torch.as_tensor(g["choose_mask"]), acts)
    np.testing.assert_allclose(p.numpy(), g["choose_prob"], rtol=1e-12, atol=1e-15)
    best = acts[p.argmax(dim=1)].numpy()
    same = (p.argmax(dim=1).numpy() == g["choose_prob"].argmax(axis=1))
    assert np.array_equal(best[same], g["choose_argmax_action"][same]) and same.mean() > 0.99


@pytest.mark.gpu
def test_k_planner_replays_the_reference_agent_trace(g):
    """ParkingAgent + RsPlanner (parking_agent.py:2-110) driven like train_HOPE_sac.py:191-225 for 64 envs x 160 steps, recorded from
    the unmodified reference: hope_planner_actions must emit the same action and executing flag at every step."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes as C
    from hope_b200 import capi
    from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
    T, n = g["planner_policy"].shape[:2]
    env = BatchedParkingEnv(n, scenes=generate_scenes(n, "Normal", 1), auto_reset=False)
    env.reset(); env.planner_reset()
    dev = env.device
    last = {"done": torch.zeros(n, dtype=torch.uint8, device=dev), "was_reset": torch.zeros(n, dtype=torch.uint8, device=dev),
            "rs_found": torch.zeros(n, dtype=torch.uint8, device=dev), "rs_nseg": torch.zeros(n, dtype=torch.uint8, device=dev),
            "rs_types": torch.full((n, 5), 255, dtype=torch.uint8, device=dev), "rs_lengths": torch.zeros((n, 5), dtype=torch.float64, device=dev)}
    st = capi.Out()
    for k, v in last.items():
        setattr(st, k, v.data_ptr())
    action = torch.zeros((n, 2), dtype=torch.float64, device=dev)
    executing = torch.zeros(n, dtype=torch.uint8, device=dev)
    n_exec = 0
    for t in range(T):
        pol = torch.as_tensor(g["planner_policy"][t], device=dev).contiguous()
        capi.check(env.lib.hope_planner_actions(env.ctx, pol.data_ptr(), C.byref(st), action.data_ptr(), executing.data_ptr(), 1.25,
                                                torch.cuda.current_stream(dev).cuda_stream), env.ctx)
        torch.cuda.synchronize()
        assert np.array_equal(executing.cpu().numpy(), g["planner_executing"][t]), t
        assert np.array_equal(action.cpu().numpy(), g["planner_action"][t]), t      # bit for bit: same divisions and subtractions
        n_exec += int(executing.sum())
        last["done"].copy_(torch.as_tensor(g["planner_done"][t])); last["rs_found"].copy_(torch.as_tensor(g["planner_found"][t]))
        last["rs_nseg"].copy_(torch.as_tensor(g["planner_nseg"][t])); last["rs_types"].copy_(torch.as_tensor(g["planner_types"][t]))
        last["rs_lengths"].copy_(torch.as_tensor(g["planner_lengths"][t]))
    assert n_exec == int(g["planner_executing"].sum()) > 3000
    env.close()
