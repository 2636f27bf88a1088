#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — feature samples of the UNMODIFIED reference's scene generator, for a distribution test.

The product's generators (hope_generate_scenes on the host, k_generate_scenes on the device; scene_gen.cu) restate
`ParkingMapNormal.reset` (parking_map_normal.py:40-494) with their own random streams, so scenes cannot be compared one by
one; tests/test_scene_distribution.py compares the DISTRIBUTIONS instead.  This script runs the reference generator
(oracle/refshim standing in for shapely) 1 500 times per level and stores, per level and case type (bay / parallel), the
sorted sample of each scalar feature computed by `features()` below — the same function the test applies to product scenes.

Writes tests/golden/scene_stats.npz.  Usage: python oracle/make_scene_stats.py [--ref /root/reference] [--n 1500]
"""
import argparse
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
FEATURES = ("n_obs", "start_x", "start_y", "start_cos", "start_sin", "dest_y", "dest_heading", "start_dest_dist", "nearest_obst_to_dest",
            "nearest_obst_to_start", "second_obst_to_dest", "obst_extent_x", "obst_y_max", "width", "height")


def _seg_dist(p, a, b):
    d = b - a
    l2 = float(d @ d)
    t = 0.0 if l2 == 0 else min(1.0, max(0.0, float((p - a) @ d) / l2))
    return float(np.hypot(*(p - (a + t * d))))


def _ring_gap(r1, r2):
    """smallest vertex-to-edge distance between two rings that do not cross"""
    best = math.inf
    for ra, rb in ((r1, r2), (r2, r1)):
        for p in ra:
            for k in range(len(rb)):
                best = min(best, _seg_dist(p, rb[k], rb[(k + 1) % len(rb)]))
    return best


def _box(pose):
    c, s = math.cos(pose[2]), math.sin(pose[2])
    return np.array([(c * x - s * y + pose[0], s * x + c * y + pose[1]) for x, y in ((-0.93, -0.97), (3.76, -0.97), (3.76, 0.97), (-0.93, 0.97))])


def features(start, dest, bounds, obs, nverts):
    """scalar features of one scene (arrays as in hope_set_scene_pool): what the distribution test compares"""
    rings = [np.asarray(obs[k][:nv], dtype=np.float64) for k, nv in enumerate(nverts) if nv]
    db, sb = _box(dest), _box(start)
    gaps_d = sorted(_ring_gap(db, r) for r in rings)
    gaps_s = sorted(_ring_gap(sb, r) for r in rings)
    allv = np.concatenate(rings, axis=0)
    return dict(n_obs=float(len(rings)), start_x=float(start[0]), start_y=float(start[1]), start_cos=math.cos(start[2]), start_sin=math.sin(start[2]),
                dest_y=float(dest[1]), dest_heading=float(dest[2]), start_dest_dist=float(np.hypot(start[0] - dest[0], start[1] - dest[1])),
                nearest_obst_to_dest=gaps_d[0], nearest_obst_to_start=gaps_s[0], second_obst_to_dest=gaps_d[1] if len(gaps_d) > 1 else gaps_d[0],
                obst_extent_x=float(allv[:, 0].max() - allv[:, 0].min()), obst_y_max=float(allv[:, 1].max()),
                width=float(bounds[1] - bounds[0]), height=float(bounds[3] - bounds[2]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--n", type=int, default=1500)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "scene_stats.npz"))
    args = ap.parse_args()
    from oracle.make_golden import _import_reference, scene_arrays
    mods = _import_reference(args.ref)
    pmn = mods[4]
    warnings.filterwarnings("ignore")
    out = {}
    for level in ("Normal", "Complex", "Extrem"):
        m = pmn.ParkingMapNormal(level)
        rows = {0: [], 1: []}
        for i in range(args.n):
            np.random.seed(700000 + i)
            m.reset(None, None)
            s, d, b, o, nv = scene_arrays(m)
            rows[int(m.case_id)].append(features(s, d, b, o, nv))
        out[f"{level}_bay_fraction"] = np.float64(len(rows[0]) / args.n)
        for case, name in ((0, "bay"), (1, "parallel")):
            for f in FEATURES:
                out[f"{level}_{name}_{f}"] = np.sort(np.array([r[f] for r in rows[case]]))
        print(level, "bay", len(rows[0]), "parallel", len(rows[1]), "n_obs mean", np.mean([r["n_obs"] for c in rows.values() for r in c]))
    np.savez_compressed(args.out, **out)


if __name__ == "__main__":
    main()
