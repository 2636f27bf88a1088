"""The constants of the reference's `src/configs.py` that the ParkingEnv path reads (SURVEY.md §5 "config / flags").

The reference star-imports one module of constants everywhere (`from configs import *`), so an edit to configs.py changes
the env.  The backend keeps that contract: `load()` returns the values of the caller's own `configs` module when one is
importable (the reference's scripts run from `src/`, so it is — `configs` is then usually in `sys.modules` already), and the
reference's shipped defaults (configs.py:13-115, 180-187) otherwise.  Every consumer in this package (`tables.py`, the
`compat/env` facade, `BatchedParkingEnv(config=...)`) takes its numbers from the object returned here, not from literals.

What the CUDA build fixes at compile time is checked, not silently ignored: `LIDAR_NUM` (120 beams), the 42-entry discrete
action grid and the 10-step mask horizon size the kernels' shared-memory arrays (include/hope_b200.h), and the image
stage's geometry (WIN 500, OBS 256, K 12, TRAJ_RENDER_LEN 20) is baked into k_render; `validate()` raises a HopeError that
names the constant when configs.py asks for something else.
"""
import importlib
import math
import sys
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np

# configs.py:13-115, 180-187 as shipped (reference commit 2accab9)
_WHEEL_BASE, _FRONT_HANG, _REAR_HANG, _WIDTH = 2.8, 0.96, 0.93, 1.94
_LENGTH = _WHEEL_BASE + _FRONT_HANG + _REAR_HANG
_NAMES = ("WHEEL_BASE", "FRONT_HANG", "REAR_HANG", "LENGTH", "WIDTH", "VALID_SPEED", "VALID_STEER", "NUM_STEP", "STEP_LENGTH",
          "MAP_LEVEL", "MIN_PARK_LOT_LEN_DICT", "MAX_PARK_LOT_LEN_DICT", "MIN_PARK_LOT_WIDTH_DICT", "MAX_PARK_LOT_WIDTH_DICT",
          "PARA_PARK_WALL_DIST_DICT", "BAY_PARK_WALL_DIST_DICT", "N_OBSTACLE_DICT", "MIN_DIST_TO_OBST", "MAX_DRIVE_DISTANCE",
          "DROUP_OUT_OBST", "ENV_COLLIDE", "BG_COLOR", "START_COLOR", "DEST_COLOR", "OBSTACLE_COLOR", "TRAJ_COLOR_HIGH",
          "TRAJ_COLOR_LOW", "TRAJ_RENDER_LEN", "OBS_W", "OBS_H", "WIN_W", "WIN_H", "LIDAR_RANGE", "LIDAR_NUM", "FPS",
          "TOLERANT_TIME", "USE_LIDAR", "USE_IMG", "USE_ACTION_MASK", "MAX_DIST_TO_DEST", "K", "RS_MAX_DIST", "RENDER_TRAJ",
          "PRECISION", "step_speed", "REWARD_RATIO", "REWARD_WEIGHT", "COLOR_POOL")


def defaults():
    c = SimpleNamespace(
        WHEEL_BASE=_WHEEL_BASE, FRONT_HANG=_FRONT_HANG, REAR_HANG=_REAR_HANG, LENGTH=_LENGTH, WIDTH=_WIDTH,
        VALID_SPEED=[-2.5, 2.5], VALID_STEER=[-0.75, 0.75], NUM_STEP=10, STEP_LENGTH=5e-2, MAP_LEVEL="Normal",
        MIN_PARK_LOT_LEN_DICT={"Extrem": _LENGTH + 0.6, "Complex": _LENGTH + 0.9, "Normal": _LENGTH * 1.25},
        MAX_PARK_LOT_LEN_DICT={"Extrem": _LENGTH + 0.9, "Complex": _LENGTH * 1.25, "Normal": _LENGTH * 1.25 + 0.5},
        MIN_PARK_LOT_WIDTH_DICT={"Complex": _WIDTH + 0.4, "Normal": _WIDTH + 0.85},
        MAX_PARK_LOT_WIDTH_DICT={"Complex": _WIDTH + 0.85, "Normal": _WIDTH + 1.2},
        PARA_PARK_WALL_DIST_DICT={"Extrem": 3.5, "Complex": 4.0, "Normal": 4.5},
        BAY_PARK_WALL_DIST_DICT={"Complex": 6.0, "Normal": 7.0},
        N_OBSTACLE_DICT={"Extrem": 8, "Complex": 5, "Normal": 3},
        MIN_DIST_TO_OBST=0.1, MAX_DRIVE_DISTANCE=15.0, DROUP_OUT_OBST=0.0, ENV_COLLIDE=False,
        BG_COLOR=(255, 255, 255, 255), START_COLOR=(100, 149, 237, 255), DEST_COLOR=(69, 139, 0, 255),
        OBSTACLE_COLOR=(150, 150, 150, 255), TRAJ_COLOR_HIGH=(10, 10, 200, 255), TRAJ_COLOR_LOW=(10, 10, 10, 255),
        TRAJ_RENDER_LEN=20, OBS_W=256, OBS_H=256, WIN_W=500, WIN_H=500, LIDAR_RANGE=10.0, LIDAR_NUM=120, FPS=100,
        TOLERANT_TIME=200, USE_LIDAR=True, USE_IMG=True, USE_ACTION_MASK=True, MAX_DIST_TO_DEST=20, K=12, RS_MAX_DIST=10,
        RENDER_TRAJ=True, PRECISION=10, step_speed=1, REWARD_RATIO=0.1,
        REWARD_WEIGHT=OrderedDict([("time_cost", 1), ("rs_dist_reward", 0), ("dist_reward", 5), ("angle_reward", 0),
                                   ("box_union_reward", 10)]),
        COLOR_POOL=[(30, 144, 255, 255), (255, 127, 80, 255), (255, 215, 0, 255)],
        source="reference defaults (configs.py:13-115, 180-187)")
    return _derive(c)


def _ring_coords(ring):
    """corner list of configs.VehicleBox (a shapely LinearRing, configs.py:20-24) without importing shapely"""
    pts = [(float(p[0]), float(p[1])) for p in ring.coords]
    if len(pts) > 1 and pts[0] == pts[-1]:
        pts = pts[:-1]
    return pts


def _derive(c, module=None):
    box = None
    if module is not None and hasattr(module, "VehicleBox"):
        try:
            box = _ring_coords(module.VehicleBox)
        except Exception:  # an exotic geometry object: fall back to the formula of configs.py:20-24
            box = None
    if box is None:
        box = [(-c.REAR_HANG, -c.WIDTH / 2), (c.FRONT_HANG + c.WHEEL_BASE, -c.WIDTH / 2),
               (c.FRONT_HANG + c.WHEEL_BASE, c.WIDTH / 2), (-c.REAR_HANG, c.WIDTH / 2)]
    c.VEHICLE_BOX = np.array(box, dtype=np.float64)
    steer = np.arange(c.VALID_STEER[-1], -(c.VALID_STEER[-1] + c.VALID_STEER[-1] / c.PRECISION), -c.VALID_STEER[-1] / c.PRECISION)
    c.discrete_actions = [[s, c.step_speed] for s in steer] + [[s, -c.step_speed] for s in steer]  # configs.py:108-115
    c.N_DISCRETE_ACTION = len(c.discrete_actions)
    return c


def from_module(module):
    """Snapshot a `configs`-like module (anything with the reference's constant names); missing names take the defaults."""
    c = defaults()
    for name in _NAMES:
        if hasattr(module, name):
            setattr(c, name, getattr(module, name))
    c.source = getattr(module, "__file__", None) or repr(module)
    return _derive(c, module)


_CACHE = {}


def load(refresh=False):
    """The caller's `configs` when importable and recognisably the reference's (it defines WHEEL_BASE and LIDAR_NUM), else
    the shipped defaults.  Cached per module object: call `load(refresh=True)` after editing a live module."""
    mod = sys.modules.get("configs")
    if mod is None:
        try:
            mod = importlib.import_module("configs")
        except Exception:  # not on the path, or its own imports (shapely, torch) are missing
            mod = None
    if mod is not None and not (hasattr(mod, "WHEEL_BASE") and hasattr(mod, "LIDAR_NUM")):
        mod = None  # some unrelated module that happens to be called `configs`
    key = id(mod)
    if refresh or key not in _CACHE:
        _CACHE.clear()
        _CACHE[key] = from_module(mod) if mod is not None else defaults()
    return _CACHE[key]


def validate(c):
    """Raise when configs.py asks for something the compiled kernels cannot do (sizes fixed in include/hope_b200.h)."""
    from .capi import HopeError, N_ACTION, N_LIDAR, N_MASK_ITER
    if int(c.LIDAR_NUM) != N_LIDAR:
        raise HopeError(f"configs.LIDAR_NUM = {c.LIDAR_NUM}: the kernels are compiled for HOPE_N_LIDAR = {N_LIDAR} beams (include/hope_b200.h)")
    if int(c.N_DISCRETE_ACTION) != N_ACTION:
        raise HopeError(f"configs.PRECISION / step_speed give {c.N_DISCRETE_ACTION} discrete actions: the kernels are compiled for "
                        f"HOPE_N_ACTION = {N_ACTION} (include/hope_b200.h)")
    if int(c.NUM_STEP) > 255 or int(c.NUM_STEP) < 1:
        raise HopeError(f"configs.NUM_STEP = {c.NUM_STEP} is outside 1..255 (substep counts travel as uint8)")
    assert N_MASK_ITER == 10  # action_mask.py:9 n_iter, not a configs.py constant
    return c


def image_supported(c):
    """k_render bakes the raster geometry of car_parking_base.py:301-350 in: only the shipped values are available."""
    return (int(c.WIN_W), int(c.WIN_H), int(c.OBS_W), int(c.OBS_H), int(c.K)) == (500, 500, 256, 256, 12) and 0 <= traj_render_len(c) <= 20


def traj_render_len(c):
    """how many trajectory boxes _render draws (car_parking_base.py:315-316): TRAJ_RENDER_LEN, none with RENDER_TRAJ off"""
    return int(c.TRAJ_RENDER_LEN) if c.RENDER_TRAJ else 0


def step_params(c):
    """dict of hope_params fields (include/hope_b200.h) from a config snapshot"""
    rw = c.REWARD_WEIGHT
    order = ("time_cost", "rs_dist_reward", "dist_reward", "angle_reward", "box_union_reward")  # car_parking_base.py:285-289
    return dict(
        wheel_base=float(c.WHEEL_BASE), box_x=[float(v) for v in c.VEHICLE_BOX[:, 0]], box_y=[float(v) for v in c.VEHICLE_BOX[:, 1]],
        valid_speed=[float(v) for v in c.VALID_SPEED], valid_steer=[float(v) for v in c.VALID_STEER], num_step=int(c.NUM_STEP),
        step_length=float(c.STEP_LENGTH), lidar_range=float(c.LIDAR_RANGE), tolerant_time=int(c.TOLERANT_TIME),
        rs_max_dist=float(c.RS_MAX_DIST), reward_weight=[float(rw[k]) for k in order], reward_ratio=float(c.REWARD_RATIO),
        env_collide=1 if c.ENV_COLLIDE else 0)


def palette(c):
    """(25, 3) uint8 colours in hope_set_palette order (configs.py:26-30, 80-88)"""
    traj = np.linspace(np.array(c.TRAJ_COLOR_LOW), np.array(c.TRAJ_COLOR_HIGH), int(c.TRAJ_RENDER_LEN), endpoint=True, dtype=np.uint8)
    rows = [c.BG_COLOR, c.OBSTACLE_COLOR, c.START_COLOR, c.DEST_COLOR, c.COLOR_POOL[0]] + [tuple(t) for t in traj[:20]]
    rows += [c.BG_COLOR] * (25 - len(rows))  # a shorter TRAJ_RENDER_LEN leaves the tail unused (hope_set_render_traj)
    return np.array([r[:3] for r in rows], dtype=np.uint8)


def max_curvature(c):
    return math.tan(c.VALID_STEER[-1]) / c.WHEEL_BASE  # car_parking_base.py:422
