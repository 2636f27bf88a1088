#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — records tests/golden/adversarial_observe.npz from the UNMODIFIED reference.

`LidarSimlator.get_observation` (env/lidar_simulator.py:31-135) and `ActionMask.get_steps` (model/action_mask.py:166-196)
of the reference, called directly on the boundary-case scenes of oracle/adversarial.py (vertices exactly on beam
directions, edges collinear with a beam, axis-aligned edges in the ego frame, vertices at exactly the 10 m range, slivers,
obstacles touching the ego position).  Same import recipe and stand-ins as oracle/make_golden.py (shapely is the restated
subset of oracle/refshim: `affine_transform` and the ring-to-point distance of the 10 m obstacle filter come from there).
Runs in the build container only.   Usage:  python oracle/make_adversarial_golden.py [--n 600]
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from adversarial import adversarial_scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "..", "tests", "golden", "adversarial_observe.npz"))
    ap.add_argument("--n", type=int, default=600)
    ap.add_argument("--seed", type=int, default=2024)
    args = ap.parse_args()
    out = os.path.abspath(args.out)
    cpb, wrap, vehicle, rs, pmn, configs = mg._import_reference(args.ref)
    from env.lidar_simulator import LidarSimlator
    from model.action_mask import ActionMask
    from shapely.geometry import LinearRing
    start, obs, nverts = adversarial_scenes(np.random.default_rng(args.seed), args.n)
    lidar_sim, mask_maker = LidarSimlator(configs.LIDAR_RANGE, configs.LIDAR_NUM), ActionMask()
    lidar, mask = np.zeros((args.n, 120)), np.zeros((args.n, 42))
    for i in range(args.n):
        rings = [LinearRing([tuple(v) for v in obs[i, k, :nverts[i, k]]]) for k in range(16) if nverts[i, k]]
        lidar[i] = lidar_sim.get_observation(vehicle.State(list(start[i])), rings)
        mask[i] = mask_maker.get_steps(lidar[i])
    np.savez_compressed(out, seed=args.seed, start=start, obs=obs, nverts=nverts, lidar=lidar, mask=mask)
    print("wrote", out, os.path.getsize(out), "bytes; beams with a hit:", int((lidar < lidar.max(axis=0) - 1e-9).sum()))


if __name__ == "__main__":
    main()
