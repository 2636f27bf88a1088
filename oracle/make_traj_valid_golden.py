#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — records tests/golden/traj_valid.npz from the UNMODIFIED reference.

Every call of `CarParking.is_traj_valid` (car_parking_base.py:452-534) made by `find_rs_path` (:413-450) during random
and RS-following episodes is recorded as: the scene, the ego pose the search started from, WHICH word of
`rsCurve.calc_all_paths` was sampled (index in the list that function returns, reeds_shepp.py:35-54), the number of
samples, and the verdict.  tests/test_rs_check_host.py replays these through the product's own k_rs_check code
(hope_b200/csrc/rs_check.cuh compiled with g++ on a 32-fiber warp emulation).

Runs in the build container only (needs /root/reference; same import recipe and stand-ins as oracle/make_golden.py).
Usage:  python oracle/make_traj_valid_golden.py [--ref /root/reference] [--out tests/golden/traj_valid.npz]
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def record(mods, level, n_episodes, seed, follow_rs, max_steps=120):
    cpb, wrap, vehicle, rs, pmn, configs = mods
    raw = cpb.CarParking(render_mode="rgb_array", fps=100, verbose=False,
                         use_lidar_observation=True, use_img_observation=False, use_action_mask=True)
    env = wrap.CarParkingWrapper(raw)
    state = {"paths": [], "calls": []}
    orig_all, orig_valid = rs.calc_all_paths, raw.is_traj_valid

    def all_spy(*a, **k):
        r = orig_all(*a, **k)
        state["paths"] = r
        return r

    def valid_spy(traj):
        verdict = orig_valid(traj)
        xs = [t[0] for t in traj]
        which = [i for i, p in enumerate(state["paths"]) if len(p.x) == len(xs) and list(p.x) == xs]
        st = raw.vehicle.state
        state["calls"].append((st.loc.x, st.loc.y, st.heading, which[0] if len(which) == 1 else -1, len(traj), bool(verdict)))
        return verdict

    rs.calc_all_paths, raw.is_traj_valid = all_spy, valid_spy
    scenes, calls = [], []
    for ep in range(n_episodes):
        np.random.seed(seed + ep)
        env.reset(None, None, level)
        scenes.append(mg.scene_arrays(raw.map))
        rng = np.random.default_rng(seed + 1000 * (ep + 1))
        queue = []
        for _ in range(max_steps):
            a = rng.uniform(-1.0, 1.0, size=2)
            if queue:
                a = np.array(queue.pop(0), dtype=np.float64)
            state["calls"] = []
            _, _, done, info = env.step(a)
            calls.extend((len(scenes) - 1,) + c for c in state["calls"])
            p = info["path_to_dest"]
            if follow_rs and p is not None and not queue:
                _, t, l = mg.encode_path(p)
                queue = mg.plan_actions(t, l, 1.25)
            if done:
                break
    rs.calc_all_paths, raw.is_traj_valid = orig_all, orig_valid
    return scenes, calls


def record_far(mods, level, n_scenes, seed, poses_per_scene=6):
    """Words longer than one 256-sample chunk: `find_rs_path` called directly (the 10 m gate of car_parking_base.py:293-294
    sits outside it) from poses 15-40 m away from the slot, with the map bounds widened so that not every word leaves
    the map.  Same spies as above; the scene recorded is the one with the widened bounds."""
    cpb, wrap, vehicle, rs, pmn, configs = mods
    raw = cpb.CarParking(render_mode="rgb_array", fps=100, verbose=False,
                         use_lidar_observation=True, use_img_observation=False, use_action_mask=True)
    env = wrap.CarParkingWrapper(raw)
    state = {"paths": [], "calls": []}
    orig_all, orig_valid = rs.calc_all_paths, raw.is_traj_valid

    def all_spy(*a, **k):
        r = orig_all(*a, **k)
        state["paths"] = r
        return r

    def valid_spy(traj):
        verdict = orig_valid(traj)
        xs = [t[0] for t in traj]
        which = [i for i, p in enumerate(state["paths"]) if len(p.x) == len(xs) and list(p.x) == xs]
        st = raw.vehicle.state
        state["calls"].append((st.loc.x, st.loc.y, st.heading, which[0] if len(which) == 1 else -1, len(traj), bool(verdict)))
        return verdict

    rs.calc_all_paths, raw.is_traj_valid = all_spy, valid_spy
    scenes, calls = [], []
    rng = np.random.default_rng(seed)
    for k in range(n_scenes):
        np.random.seed(seed + k)
        env.reset(None, None, level)
        m = raw.map
        m.xmin -= 45.0; m.xmax += 45.0; m.ymin -= 45.0; m.ymax += 45.0
        scenes.append(mg.scene_arrays(m))
        dx, dy, _ = m.dest.get_pos()
        for _ in range(poses_per_scene):
            r, th = rng.uniform(15.0, 40.0), rng.uniform(-np.pi, np.pi)
            raw.vehicle.state = vehicle.State([dx + r * np.cos(th), dy + r * np.sin(th), rng.uniform(-np.pi, np.pi), 0.0, 0.0])
            state["calls"] = []
            raw.find_rs_path(None)
            calls.extend((len(scenes) - 1,) + c for c in state["calls"])
    rs.calc_all_paths, raw.is_traj_valid = orig_all, orig_valid
    return scenes, calls


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden", "traj_valid.npz"))
    ap.add_argument("--episodes", type=int, default=8)
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    mods = mg._import_reference(args.ref)
    all_scenes, all_calls = [], []
    for li, level in enumerate(("Normal", "Complex", "Extrem")):
        for follow in (False, True):
            scenes, calls = record(mods, level, args.episodes, 7000 + 100 * li + (50 if follow else 0), follow)
            base = len(all_scenes)
            all_scenes.extend(scenes)
            all_calls.extend((c[0] + base,) + c[1:] for c in calls)
            print(level, "follow" if follow else "random", len(calls), "calls,", sum(c[-1] for c in calls), "valid", flush=True)
    calls = [c for c in all_calls if c[4] >= 0]  # the sampled word was identified unambiguously
    print("kept", len(calls), "of", len(all_calls))
    # keep every valid (True) verdict and a strided subset of the invalid ones
    keep = [c for c in calls if c[-1]] + [c for c in calls if not c[-1]][::3]
    for li, level in enumerate(("Normal", "Complex")):  # long words from far poses, all kept
        scenes, calls = record_far(mods, level, 12, 9000 + 100 * li)
        base = len(all_scenes)
        all_scenes.extend(scenes)
        far = [(c[0] + base,) + c[1:] for c in calls if c[4] >= 0]
        keep.extend(far)
        print(level, "far", len(far), "calls,", sum(c[-1] for c in far), "valid, longest", max(c[5] for c in far), flush=True)
    keep.sort(key=lambda c: c[0])
    np.savez_compressed(
        out,
        scene_start=np.array([s[0] for s in all_scenes]), scene_dest=np.array([s[1] for s in all_scenes]),
        scene_bounds=np.array([s[2] for s in all_scenes]), scene_obs=np.array([s[3] for s in all_scenes]),
        scene_nverts=np.array([s[4] for s in all_scenes]),
        call_scene=np.array([c[0] for c in keep], dtype=np.int32), call_pose=np.array([c[1:4] for c in keep], dtype=np.float64),
        call_word=np.array([c[4] for c in keep], dtype=np.int32), call_T=np.array([c[5] for c in keep], dtype=np.int32),
        call_valid=np.array([c[6] for c in keep], dtype=np.uint8))
    print("wrote", out, len(keep), "calls,", sum(c[-1] for c in keep), "valid;", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
