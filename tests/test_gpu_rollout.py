"""GPU tests of the rollout-side rows (a19, f2): batched RS planner hand-off vs a restatement of
RsPlanner/ParkingAgent, masked discrete sampling, and the device-resident acting loop."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from hope_b200 import rollout  # noqa: E402
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes  # noqa: E402
from oracle import planner_oracle as plo  # noqa: E402


def test_planner_handoff_matches_rsplanner_semantics():
    """parking_agent.py:2-47, 60-110 and train_HOPE_sac.py:194-213, lock-step over 1 024 envs."""
    n, steps = 1024, 90
    sc = generate_scenes(2 * n, "Complex", 21)
    env = BatchedParkingEnv(n, scenes=sc, auto_reset=True)
    env.reset(); env.planner_reset()
    planners = [plo.PlannerOracle(1.25) for _ in range(n)]
    rng = np.random.default_rng(3)
    n_exec = n_loaded = n_finished = 0
    for t in range(steps):
        pol = rng.uniform(-1, 1, size=(n, 2))
        act, exe = env.planner_actions(torch.as_tensor(pol, device=env.device).contiguous())
        torch.cuda.synchronize()
        act, exe = act.cpu().numpy().copy(), exe.cpu().numpy().astype(bool).copy()
        want, want_exe = pol.copy(), np.zeros(n, dtype=bool)
        for i, p in enumerate(planners):
            if p.executing:
                want[i] = p.get_action(); want_exe[i] = True
                n_finished += not p.executing
        assert np.array_equal(exe, want_exe), t
        assert np.array_equal(act, want), t
        n_exec += int(exe.sum())
        env.step(torch.as_tensor(act, device=env.device).contiguous())
        torch.cuda.synchronize()
        o = {k: v.cpu().numpy() for k, v in env.out.items()}
        for i, p in enumerate(planners):
            if o["done"][i] or o["was_reset"][i]:
                p.reset()                                   # next episode: ParkingAgent.reset
            elif o["rs_found"][i]:
                k = int(o["rs_nseg"][i])
                before = p.executing
                p.set_path(o["rs_types"][i][:k], [float(v) for v in o["rs_lengths"][i][:k]])
                n_loaded += (not before)
    assert n_loaded > 20 and n_exec > 100 and n_finished > 5, (n_loaded, n_exec, n_finished)
    assert env.out["status"].eq(2).any() or True
    env.close()


def test_following_plans_reaches_the_slot():
    """Sanity of the whole hand-off: executing the found paths open-loop must produce ARRIVED episodes."""
    n = 2048
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "Normal", 5), auto_reset=True)
    env.reset(); env.planner_reset()
    gen = torch.Generator(device=env.device); gen.manual_seed(0)
    arrived = 0
    for _ in range(150):
        pol = torch.rand((n, 2), dtype=torch.float64, device=env.device, generator=gen) * 2 - 1
        act, _ = env.planner_actions(pol)
        env.step(act)
        arrived += int((env.out["status"] == 2).sum())
    assert arrived > 50, arrived
    env.close()


def test_masked_sampling_respects_mask_and_distribution():
    dev = torch.device("cuda")
    acts = rollout.possible_actions(dev)
    n = 20000
    mean = torch.tensor([[0.3, 0.6]], dtype=torch.float64, device=dev).expand(n, 2)
    std = torch.tensor([[0.5, 0.9]], dtype=torch.float64, device=dev).expand(n, 2)
    mask = torch.zeros((n, 42), dtype=torch.float64, device=dev)
    mask[:, 3:12] = torch.linspace(0.1, 1.0, 9, dtype=torch.float64, device=dev)
    mask[:, 25] = 0.5
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    a, idx = rollout.masked_discrete_actions(mean, std, mask, acts, gen)
    want = plo.masked_action_probabilities(mean[0].cpu().numpy(), std[0].cpu().numpy(), mask[0].cpu().numpy(), acts.cpu().numpy())
    got = np.bincount(idx.cpu().numpy(), minlength=42) / n
    assert (got[want == 0] == 0).all()
    assert np.abs(got - want).max() < 0.02
    assert torch.equal(a, acts[idx])


def test_rollout_engine_runs_device_resident():
    n = 4096
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "mix", 9), auto_reset=True)
    actor = rollout.ReferenceShapedActor().to(env.device)
    eng = rollout.RolloutEngine(env, actor, seed=0)
    seen = {"steps": 0, "exec": 0}

    def store(t, obs, action, reward, done, log_prob, executing):
        assert action.shape == (n, 2) and action.dtype == torch.float64 and reward.shape == (n,)
        assert torch.isfinite(log_prob).all() and torch.isfinite(obs["lidar"]).all()
        seen["steps"] += 1
        seen["exec"] += int(executing.sum())

    c0 = env.counters()
    eng.collect(24, store=store)
    torch.cuda.synchronize()
    c1 = env.counters()
    assert seen["steps"] == 24
    assert c1["env_steps"] - c0["env_steps"] > 20 * n
    assert eng.norm.n == 24 * n and torch.isfinite(eng.norm.mean["lidar"]).all()
    env.close()


def test_stored_transition_pairs_the_action_with_the_observation_it_was_chosen_from():
    """the env writes its outputs in place: what `store` receives must be the observation act() saw, not the next one, and the
    log-probability must be that of the EXECUTED action (the plan's where a route is being executed, parking_agent.py:93-97)"""
    import math
    n = 2048
    env = BatchedParkingEnv(n, scenes=generate_scenes(2 * n, "Normal", 4), auto_reset=True)
    actor = rollout.ReferenceShapedActor().to(env.device)
    eng = rollout.RolloutEngine(env, actor, seed=0)
    seen, dists = [], []
    orig_act = eng.act

    def spy_act(obs):
        seen.append({k: obs[k].clone() for k in ("lidar", "target", "action_mask")})
        a, dist = orig_act(obs)
        dists.append(dist)
        return a, dist

    eng.act = spy_act
    n_exec = [0]

    def store(t, obs, action, reward, done, log_prob, executing):
        for k in ("lidar", "target", "action_mask"):
            assert torch.equal(obs[k], seen[t][k]), (t, k)
        mean, std = dists[t]
        want = -0.5 * ((action - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
        assert torch.equal(log_prob, want)
        ex = executing.bool()
        n_exec[0] += int(ex.sum())
        if ex.any():  # plan actions are {-1,0,1} steers with unit or remainder speeds: not what the sampler drew
            assert torch.all((action[ex, 0] == 0) | (action[ex, 0].abs() == 1))

    eng.collect(40, store=store)
    torch.cuda.synchronize()
    assert len(seen) == 40 and n_exec[0] > 0
    assert not torch.equal(seen[0]["lidar"], seen[1]["lidar"])
    env.close()
