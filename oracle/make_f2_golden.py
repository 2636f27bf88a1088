#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — pins the rollout-side helpers (scope row f2) on the UNMODIFIED reference.

Runs in the build container only (needs /root/reference).  Imported as they lie: `model/agent/parking_agent.py`
(RsPlanner, ParkingAgent — no third-party imports), `model/state_norm.py`, `model/replay_memory.py` (numpy only) and
`model/action_mask.py` (ActionMask.choose_action; its module imports shapely, for which oracle/refshim stands in).

Writes tests/golden/f2_helpers.npz:
  planner_*   an event trace of ParkingAgent + RsPlanner for 64 envs x 160 steps driven like the inner loop of
              train_HOPE_sac.py:191-225: per step the action the agent emits (the planner's when a route is being executed,
              else the policy's), then the events of that step (a Reeds-Shepp path handed over / the episode ending).
              Paths include the edge cases of set_rs_path (:12-41): segments of exactly one step, shorter than 1e-3 steps,
              long negative segments, five segments.  This is what hope_planner_actions (k_planner) must reproduce.
  norm_*      StateNorm (state_norm.py:25-46) fed 1 536 lidar / target observations one at a time with update=True:
              final mean / std / n and the normalised value of a probe observation with update=False.
  replay_*    ReplayMemory (replay_memory.py:6-50) with capacity 96 after 250 pushes of one env's transitions: the transition
              ids it still holds, in deque order, and get_items() of every slot (done flags and whether next_state exists).
  choose_*    ActionMask.choose_action's probabilities (action_mask.py:199-227) for 200 (mean, std, mask) triples, captured by
              substituting np.random.choice for the duration of the call.

Usage: python oracle/make_f2_golden.py [--ref /root/reference]
"""
import argparse
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LETTER = {0: "S", 1: "L", 2: "R"}


class _Path(object):  # what info['path_to_dest'] carries (reeds_shepp.py:15-32): the planner reads .ctypes and .lengths
    def __init__(self, types, lengths):
        self.ctypes = [LETTER[int(t)] for t in types]
        self.lengths = [float(v) for v in lengths]


class _Policy(object):
    """stands in for the RL agent behind ParkingAgent: get_action returns the scripted policy output"""

    def __init__(self):
        self.next = None

    def get_action(self, obs):
        return self.next, None

    def get_log_prob(self, obs, action):
        return 0.0


def record_planner(pa, rng, n_env=64, n_step=160, step_ratio=1.25):
    agents = []
    for _ in range(n_env):
        pol = _Policy()
        agents.append((pa.ParkingAgent(pol, pa.RsPlanner(step_ratio)), pol))
    rec = dict(policy=np.zeros((n_step, n_env, 2)), action=np.zeros((n_step, n_env, 2)), executing=np.zeros((n_step, n_env), dtype=np.uint8),
               found=np.zeros((n_step, n_env), dtype=np.uint8), nseg=np.zeros((n_step, n_env), dtype=np.uint8),
               types=np.full((n_step, n_env, 5), 255, dtype=np.uint8), lengths=np.zeros((n_step, n_env, 5)),
               done=np.zeros((n_step, n_env), dtype=np.uint8))
    special = [1.25, -1.25, 1.25e-3 * 0.5, -1.25e-3 * 0.5, 2.5, -3.75, 1.2499999999999998, 0.0012500000000000002]
    for t in range(n_step):
        for e, (agent, pol) in enumerate(agents):
            p = rng.uniform(-1, 1, size=2)
            pol.next = p.copy()
            rec["policy"][t, e] = p
            rec["executing"][t, e] = 1 if agent.executing_rs else 0
            a, _ = agent.get_action(None)                      # train_HOPE_sac.py:198 -> parking_agent.py:99-110
            rec["action"][t, e] = a
            # events of this env step (after env.step): path handed over (:212-213) or episode over (:184 parking_agent.reset)
            u = rng.random()
            if u < 0.06:
                rec["done"][t, e] = 1
                agent.reset()
            elif u < 0.30:
                k = int(rng.integers(1, 6))
                types = rng.integers(0, 3, size=k)
                lengths = rng.uniform(-6.0, 6.0, size=k)
                for j in range(k):
                    if rng.random() < 0.25:
                        lengths[j] = special[int(rng.integers(0, len(special)))]
                steps = np.abs(lengths / step_ratio)
                if not ((steps > 1e-3) & (steps != 1.0)).any():
                    lengths[0] = 2.0  # a path whose every segment is dropped leaves the reference popping an empty list (IndexError, :44)
                rec["found"][t, e] = 1; rec["nseg"][t, e] = k
                rec["types"][t, e, :k] = types; rec["lengths"][t, e, :k] = lengths
                agent.set_planner_path(_Path(types, lengths))  # ignored while a route is being executed (:64-68)
    return {"planner_" + k: v for k, v in rec.items()}


def record_norm(sn_mod, rng, n=1536):
    shape = {"lidar": (120,), "target": (5,), "action_mask": (42,), "img": (3, 64, 64)}
    sn = sn_mod.StateNorm(shape)
    lidar = rng.uniform(0.0, 10.0, size=(n, 120)) * (rng.random(size=(n, 120)) < 0.7) + 0.25
    target = np.concatenate([rng.uniform(0, 20, size=(n, 1)), rng.uniform(-1, 1, size=(n, 4))], axis=1)
    first = None
    for i in range(n):
        obs = {"lidar": lidar[i].copy(), "target": target[i].copy(), "action_mask": np.ones(42), "img": np.zeros((3, 64, 64))}
        out = sn.state_norm(obs, update=True)
        if i == 1:
            first = {k: out[k].copy() for k in ("lidar", "target")}
    probe = {"lidar": np.linspace(0, 10, 120), "target": np.array([7.0, 0.3, -0.4, 0.9, 0.9]), "action_mask": np.ones(42), "img": np.zeros((3, 64, 64))}
    pn = sn.state_norm({k: v.copy() for k, v in probe.items()}, update=False)
    return dict(norm_lidar=lidar, norm_target=target, norm_n=np.int64(sn.n_state), norm_mean_lidar=sn.state_mean["lidar"],
                norm_mean_target=sn.state_mean["target"], norm_std_lidar=sn.state_std["lidar"], norm_std_target=sn.state_std["target"],
                norm_probe_lidar=probe["lidar"], norm_probe_target=probe["target"], norm_probe_out_lidar=pn["lidar"],
                norm_probe_out_target=pn["target"], norm_second_out_lidar=first["lidar"], norm_second_out_target=first["target"])


def record_replay(rm_mod, rng, capacity=96, pushes=250):
    mem = rm_mod.ReplayMemory(capacity, ["log_prob", "next_obs"])
    done = rng.random(size=pushes) < 0.08
    for i in range(pushes):
        state = np.array([float(i), 0.5 * i])
        mem.push((state, np.array([0.1 * i, -0.1 * i]), float(i) * 0.01, bool(done[i]), -float(i), np.array([float(i + 1), 0.5 * (i + 1)])))
    ids = np.array([int(s[0]) for s in mem.memory["state"]])
    # one slot per call: with a None among the next states numpy >= 1.24 refuses the ragged np.array() of get_items (:31-33)
    rows = [mem.get_items(np.array([k])) for k in range(len(mem))]
    b = {key: [r[key][0] for r in rows] for key in rows[0]}
    has_next = np.array([x is not None for x in b["next_state"]], dtype=np.uint8)
    next_id = np.array([int(x[0]) if x is not None else -1 for x in b["next_state"]])
    np.random.seed(5)
    sample_idx = np.random.randint(len(mem), size=64)
    return dict(replay_capacity=np.int64(capacity), replay_pushes=np.int64(pushes), replay_done=done.astype(np.uint8), replay_ids=ids,
                replay_len=np.int64(len(mem)), replay_has_next=has_next, replay_next_id=next_id,
                replay_reward=np.array(b["reward"]), replay_action=np.array(b["action"]), replay_sample_idx=sample_idx)


def record_choose(am_mod, rng, n=200):
    mask_obj = am_mod.ActionMask.__new__(am_mod.ActionMask)  # choose_action only reads self.action_space (action_mask.py:217)
    import configs
    mask_obj.action_space = configs.discrete_actions
    mean = rng.uniform(-1, 1, size=(n, 2)); std = np.exp(rng.uniform(-1.5, 0.5, size=(n, 2)))
    mask = np.round(rng.random(size=(n, 42)) * 10) / 10 * (rng.random(size=(n, 42)) < 0.6)
    mask[:, 0] = np.maximum(mask[:, 0], 0.1)  # at least one admissible action
    mask[7] = 0.01                             # the all-zero mask case (action_mask.py:182-183)
    probs = np.zeros((n, 42)); chosen = np.zeros((n, 2))
    orig = np.random.choice
    try:
        for i in range(n):
            grabbed = {}

            def fake_choice(a, p=None):
                grabbed["p"] = np.array(p)
                return int(np.argmax(p))
            np.random.choice = fake_choice
            chosen[i] = mask_obj.choose_action(mean[i], std[i], mask[i])
            probs[i] = grabbed["p"]
    finally:
        np.random.choice = orig
    return dict(choose_mean=mean, choose_std=std, choose_mask=mask, choose_prob=probs, choose_argmax_action=chosen)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "f2_helpers.npz"))
    args = ap.parse_args()
    src = os.path.join(args.ref, "src")
    sys.path.insert(0, src)
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    os.chdir(src)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import model.agent.parking_agent as pa
    import model.state_norm as sn
    import model.replay_memory as rm
    import model.action_mask as am
    for mod in (pa, sn, rm, am):
        assert mod.__file__.startswith(src), mod.__file__
    rng = np.random.default_rng(20240529)
    out = {}
    out.update(record_planner(pa, rng))
    out.update(record_norm(sn, rng))
    out.update(record_replay(rm, rng))
    out.update(record_choose(am, rng))
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "planner executing steps", int(out["planner_executing"].sum()), "paths", int(out["planner_found"].sum()),
          "replay kept", int(out["replay_len"]))


if __name__ == "__main__":
    main()
