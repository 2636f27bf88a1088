"""Dragon Lake Parking scenes (scope row f3): reader for the reference's `data/dlp.data` and the scene
preparation of `ParkingMapDLP.reset` (src/env/parking_map_dlp.py:38-101), without shapely.

`dlp.data` is a pickle of 248 cases `(start_candidates, dest, obstacles)`; the obstacles are shapely 1.x
`LinearRing`s, which pickle as class + WKB bytes of a closed LineString.  The unpickler below maps that
class to a plain holder and decodes the WKB, so neither shapely nor GEOS is needed.

Prepared scenes use the 128-ring capacity build (`BatchedParkingEnv(..., max_obs=128)`): after the
reference's own bounding-box filter a case keeps 37-125 obstacle rings.
"""
import pickle
import struct

import numpy as np

MAX_OBS_DLP = 128
_BOX = np.array([(-0.93, -0.97), (3.76, -0.97), (3.76, 0.97), (-0.93, 0.97)])  # configs.py:20-24


class _Ring(object):
    """stands in for shapely.geometry.polygon.LinearRing while unpickling"""

    def __setstate__(self, state):
        self.xy = _decode_wkb_linestring(bytes(state))


def _decode_wkb_linestring(b):
    order = "<" if b[0] == 1 else ">"
    (gtype,) = struct.unpack_from(order + "I", b, 1)
    dims = 3 if (gtype & 0x80000000 or gtype // 1000 == 1) else 2
    if (gtype & 0xFF) != 2:
        raise ValueError(f"expected a WKB LineString, got geometry type {gtype}")
    (n,) = struct.unpack_from(order + "I", b, 5)
    pts = np.frombuffer(b, dtype=np.dtype(order + "f8"), count=n * dims, offset=9).reshape(n, dims)[:, :2]
    return np.array(pts, dtype=np.float64)


# the only globals the reference's dlp.data refers to (besides the ring class): numpy scalars of a given dtype
_ALLOWED = {("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy", "dtype"),
            ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray")}


class _Unpickler(pickle.Unpickler):
    """Exact allow-list: anything else (builtins.eval, os.system, ...) is refused, so a crafted file cannot run code."""

    def find_class(self, module, name):
        if module.startswith("shapely"):
            if name == "LinearRing":
                return _Ring
            raise pickle.UnpicklingError(f"unexpected shapely class {module}.{name}")
        if (module, name) not in _ALLOWED:
            raise pickle.UnpicklingError(f"refusing to load {module}.{name}")
        return super().find_class(module, name)


def read_dlp(path):
    """-> list of cases, each dict(starts (k,3) f64, dest (3,) f64, rings [ (nv,2) f64 open vertex lists ])"""
    with open(path, "rb") as f:
        raw = _Unpickler(f).load()
    cases = []
    for case in raw:
        start, dest, obstacles = case[:3]
        starts = np.array(start, dtype=np.float64).reshape(-1, 3) if isinstance(start, list) else np.array([start], dtype=np.float64)
        rings = []
        for r in obstacles:
            xy = r.xy
            if len(xy) > 1 and np.array_equal(xy[0], xy[-1]):
                xy = xy[:-1]
            rings.append(xy.copy())
        cases.append(dict(starts=starts, dest=np.array(dest, dtype=np.float64), rings=rings))
    multi = isinstance(raw[0][0], list)  # ParkingMapDLP.__init__ decides `multi_start` on the first case (parking_map_dlp.py:33-35)
    for c in cases:
        c["multi"] = multi
    return cases


def _flip(pose):
    """ParkingMapDLP._flip_box_orientation (:113-119): same box, heading + pi"""
    c, s = np.cos(pose[2]), np.sin(pose[2])
    corners = np.stack([c * _BOX[:, 0] - s * _BOX[:, 1] + pose[0], s * _BOX[:, 0] + c * _BOX[:, 1] + pose[1]], axis=1)
    centre = np.mean(corners, axis=0)
    return np.array([2 * centre[0] - pose[0], 2 * centre[1] - pose[1], pose[2] + np.pi])


def prepare_scene(case, rng, start_index=None, flips=None, max_obs=MAX_OBS_DLP):
    """One `ParkingMapDLP.reset` (:38-86): pick and jitter a start, bounds = floor/ceil of the poses -/+ 20,
    keep the rings whose bounding box reaches into the bounds, flip dest / start with p = 0.5 each.
    `rng` is a numpy Generator, or any object with integers(lo, hi) / standard_normal(n) / random(): the facade passes
    numpy's global generator (the one the reference draws from) and the draws below come in the reference's order
    (:62 start index, :64 three jitters, :81 dest flip, :83 start flip), so the same np.random.seed gives the same scene."""
    starts = case["starts"]
    if case.get("multi", len(starts) > 1):
        k = int(rng.integers(0, len(starts))) if start_index is None else start_index
        st = starts[k] + rng.standard_normal(3) * np.array([0.05, 0.05, 0.02])
    else:
        st = starts[0].copy()
    dest = case["dest"].copy()
    xmin, xmax = np.floor(min(st[0], dest[0]) - 20), np.ceil(max(st[0], dest[0]) + 20)
    ymin, ymax = np.floor(min(st[1], dest[1]) - 20), np.ceil(max(st[1], dest[1]) + 20)
    kept = [r for r in case["rings"]
            if not (r[:, 0].max() <= xmin or r[:, 0].min() >= xmax or r[:, 1].max() <= ymin or r[:, 1].min() >= ymax)]
    if len(kept) > max_obs:
        raise ValueError(f"{len(kept)} obstacle rings after filtering, capacity {max_obs}")
    f = (rng.random() > 0.5, rng.random() > 0.5) if flips is None else flips
    if f[0]:
        dest = _flip(dest)
    if f[1]:
        st = _flip(st)
    obs = np.zeros((max_obs, 4, 2))
    nverts = np.zeros(max_obs, dtype=np.int32)
    for k, r in enumerate(kept):
        if not 3 <= len(r) <= 4:
            raise ValueError(f"ring with {len(r)} vertices")
        obs[k, :len(r)] = r
        nverts[k] = len(r)
    return dict(start=st, dest=dest, bounds=np.array([xmin, xmax, ymin, ymax]), obs=obs, nverts=nverts)


def prepare_scenes(cases, case_ids, seed=0, max_obs=MAX_OBS_DLP):
    rng = np.random.default_rng(seed)
    rows = [prepare_scene(cases[int(c) % len(cases)], rng, max_obs=max_obs) for c in case_ids]
    out = {k: np.stack([r[k] for r in rows]) for k in rows[0]}
    out["case_id"] = np.asarray(case_ids, dtype=np.int32)
    return out


def cases_from_fixture(npz):
    """cases stored by oracle/make_dlp_fixture.py (a subset of dlp.data as plain arrays)"""
    cases = []
    for j in range(len(npz["case_ids"])):
        nv = npz[f"ring_nv_{j}"]
        rings = [npz[f"rings_{j}"][k, :nv[k]].copy() for k in range(len(nv))]
        cases.append(dict(starts=npz[f"starts_{j}"].copy(), dest=npz[f"dest_{j}"].copy(), rings=rings, multi=True))
    return cases
