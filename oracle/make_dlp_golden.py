#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — pins the Dragon Lake Parking scene preparation (scope row f3) on the UNMODIFIED reference.

Runs in the build container only (needs /root/reference).  The reference's `env/parking_map_dlp.py` and
`env/map_level.py` are imported as they lie, with oracle/refshim standing in for shapely (README there), and

  1. `ParkingMapDLP.reset(case_id)` (parking_map_dlp.py:38-86) is run over the 16 cases of tests/golden/dlp_cases.npz
     (written to a temporary pickle in the layout of data/dlp.data, so that the recording can be replayed wherever the
     fixture is, without the reference tree) under `np.random.seed(s)` for a list of seeds, with and without an explicit
     case id: the recorded start / dest / bounds / kept obstacles / map_level are what `hope_b200.dlp.prepare_scene`
     and `compat/env/map_level.get_map_level` must reproduce from the same seed;
  2. `get_map_level` (map_level.py:27-112) labels all 248 cases of data/dlp.data (first start candidate, no jitter, no flips).

Writes tests/golden/dlp_reset.npz.  Usage: python oracle/make_dlp_golden.py [--ref /root/reference]
"""
import argparse
import os
import pickle
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LEVEL_CODE = {"Normal": 0, "Complex": 1, "Extrem": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "dlp_reset.npz"))
    args = ap.parse_args()
    src = os.path.join(args.ref, "src")
    sys.path.insert(0, src)
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    os.chdir(src)  # ParkingMapDLP.default['path'] is relative to src/
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    from shapely.geometry import LinearRing
    import env.parking_map_dlp as pmd
    from env.map_level import get_map_level
    from env.vehicle import State
    from env.map_base import Area

    fx = np.load(os.path.join(ROOT, "tests", "golden", "dlp_cases.npz"))
    data = []
    for j in range(len(fx["case_ids"])):
        nv = fx[f"ring_nv_{j}"]
        rings = [LinearRing([tuple(p) for p in fx[f"rings_{j}"][k, :nv[k]]]) for k in range(len(nv))]
        starts = [tuple(float(v) for v in s) for s in fx[f"starts_{j}"]]
        data.append((starts, tuple(float(v) for v in fx[f"dest_{j}"]), rings))
    with tempfile.NamedTemporaryFile(suffix=".data", delete=False) as f:
        pickle.dump(data, f)
        tmp = f.name

    m = pmd.ParkingMapDLP()          # loads ../data/dlp.data
    rec = {k: [] for k in ("seed", "case_arg", "case_id", "start", "dest", "bounds", "n_obst", "obst_sum", "level", "kept")}
    first = True
    for seed in range(40):
        for case_arg in (None, seed % 16, 16 + seed):  # random case, explicit case, explicit case beyond the list (wraps, :51-52)
            np.random.seed(1000 + seed)
            m.reset(case_arg, tmp if first else None)
            first = False
            rec["seed"].append(1000 + seed); rec["case_arg"].append(-1 if case_arg is None else case_arg); rec["case_id"].append(m.case_id)
            rec["start"].append(m.start.get_pos()); rec["dest"].append(m.dest.get_pos())
            rec["bounds"].append([m.xmin, m.xmax, m.ymin, m.ymax])
            rec["n_obst"].append(len(m.obstacles))
            rec["obst_sum"].append(float(sum(np.array(a.shape.coords)[:-1].sum() for a in m.obstacles)))
            rec["level"].append(LEVEL_CODE[m.map_level])
            ident = {id(r): k for k, r in enumerate(m.map_data[m.case_id][2])}
            kept = np.zeros(512, dtype=np.uint8)
            for a in m.obstacles:
                kept[ident[id(a.shape)]] = 1
            rec["kept"].append(kept)
    os.unlink(tmp)

    full = pmd.ParkingMapDLP()       # all 248 cases of data/dlp.data: label of (first start candidate, dest, all obstacles)
    levels = []
    for case in full.map_data:
        start, dest, obstacles = case[:3]
        s0 = start[0] if isinstance(start, list) else start
        areas = [Area(shape=o, subtype="obstacle", color=None) for o in obstacles]  # a bare LinearRing list hits an unbound local in :33-38
        levels.append(LEVEL_CODE[get_map_level(State(list(s0)), State(list(dest)), areas)])
    out = {k: np.array(v) for k, v in rec.items()}
    out["levels_all"] = np.array(levels, dtype=np.int32)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, {k: v.shape for k, v in out.items()}, "labels", np.bincount(out["levels_all"], minlength=3),
          "reset labels", np.bincount(out["level"], minlength=3))


if __name__ == "__main__":
    main()
