#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — times the UNMODIFIED reference's `CarParkingWrapper.step` in the build container.

BASELINE cfg 1: one env, level given, 200 float64 random actions per episode, image off, lidar + action
mask on, one warm-up episode, `clock.tick` is a no-op in the pygame stand-in (as shipped the reference self-limits to
100 steps/s, car_parking_base.py:409).  The reference sources are imported where they lie under /root/reference, with
oracle/refshim standing in for shapely / gym / pygame / heapdict (restated geometry: the per-step cost of the GEOS
predicates is that of the pure-Python stand-in, not of real shapely, so this number is a LOWER bound on the reference's
speed with real shapely; SURVEY.md §6 estimates the difference).

This cannot run on the GPU box (no /root/reference there); bench.py's reference arm times the C oracle port instead.
The record it writes (profiles/r01_reference_python_timing.json) is the in-container measurement that the bench line's
`cpu_baseline.sample` text quotes.

Usage:  python oracle/time_reference.py [--ref /root/reference] [--episodes 5] [--out profiles/...json]
"""
import argparse
import json
import os
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (same import recipe as the golden generator)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--episodes", type=int, default=5)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(HERE, "..", "profiles", "r01_reference_python_timing.json"))
    args = ap.parse_args()
    out_path = os.path.abspath(args.out)
    cpb, wrap, vehicle, rs, pmn, configs = mg._import_reference(args.ref)

    results = {}
    for level in ("Normal", "Complex", "Extrem"):
        raw = cpb.CarParking(render_mode="rgb_array", fps=100, verbose=False,
                             use_lidar_observation=True, use_img_observation=False, use_action_mask=True)
        env = wrap.CarParkingWrapper(raw)
        rng = np.random.default_rng(42)
        np.random.seed(42)
        n_steps, n_resets, t_step, t_reset = 0, 0, 0.0, 0.0
        for ep in range(args.episodes + 1):  # episode 0 is the warm-up
            t0 = time.perf_counter()
            env.reset(None, None, level)
            t1 = time.perf_counter()
            if ep:
                t_reset += t1 - t0
                n_resets += 1
            for _ in range(args.steps):
                a = rng.uniform(-1.0, 1.0, size=2)  # float64 (NumPy-2 promotion trap with float32)
                t0 = time.perf_counter()
                _, _, done, _ = env.step(a)
                t1 = time.perf_counter()
                if ep:
                    t_step += t1 - t0
                    n_steps += 1
                if done:
                    t0 = time.perf_counter()
                    env.reset(None, None, level)
                    t1 = time.perf_counter()
                    if ep:
                        t_reset += t1 - t0
                        n_resets += 1
        results[level] = {"env_steps": n_steps, "steps_per_s": n_steps / t_step, "ms_per_step": 1e3 * t_step / n_steps,
                          "resets": n_resets, "ms_per_reset": 1e3 * t_reset / max(1, n_resets)}
        print(level, results[level], flush=True)
    rec = {"what": "unmodified reference CarParkingWrapper.step, single process, one core, image off, lidar + action mask on",
           "where": "build container (no GPU); shapely/gym/pygame/heapdict are the oracle/refshim stand-ins",
           "cpu": platform.processor() or platform.machine(), "cores_used": 1, "host_cores": os.cpu_count(),
           "episodes_timed": args.episodes, "steps_per_episode": args.steps, "levels": results,
           "mean_steps_per_s": float(np.mean([r["steps_per_s"] for r in results.values()]))}
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
