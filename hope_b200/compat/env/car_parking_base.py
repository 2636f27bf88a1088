"""CarParking facade (car_parking_base.py:39-541 surface) over one env of the CUDA backend.

Same constructor, attributes and return values as the reference class for the lidar / target / action-mask / image
modalities.  Every constant (defaults of the constructor, action space, observation space, step parameters, colours) is
read from the caller's own `configs` module through hope_b200.refconfig, so edits to the reference's configs.py keep
taking effect.  Differences, all by construction of the backend:
  * generated scenes (levels Normal / Complex / Extrem) come from hope_generate_scenes (same construction and
    distributions as parking_map_normal.py, own per-scene RNG stream); the stream's seed is drawn from numpy's global
    generator at every reset, so `np.random.seed(s)` (train_HOPE_sac.py:137) still makes a run reproducible;
  * Dragon Lake Parking scenes (level 'dlp') are prepared exactly like ParkingMapDLP.reset (parking_map_dlp.py:38-86),
    drawing from numpy's global generator in the reference's order: the same seed gives the same case, start and flips;
  * no pygame window, no clock.tick (the reference sleeps to <= fps steps/s, :409); `render()` returns the current
    observation without opening anything;
  * the image (scope row f1) is rendered by k_render as uint8 and divided by 255.0 here, which is exactly what
    Obs_Processor.process_img returns (observation_processor.py:13-17).
"""
from collections import OrderedDict

import numpy as np

from hope_b200 import capi, refconfig
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
from env.map_base import Area
from env.vehicle import Ring, State, Status, Vehicle

_C = refconfig.load()
VALID_SPEED, VALID_STEER = list(_C.VALID_SPEED), list(_C.VALID_STEER)
NUM_STEP, STEP_LENGTH = _C.NUM_STEP, _C.STEP_LENGTH
LIDAR_NUM, LIDAR_RANGE, N_DISCRETE_ACTION, MAX_DIST_TO_DEST = _C.LIDAR_NUM, _C.LIDAR_RANGE, _C.N_DISCRETE_ACTION, _C.MAX_DIST_TO_DEST
FPS, USE_LIDAR, USE_IMG, USE_ACTION_MASK, MAP_LEVEL = _C.FPS, _C.USE_LIDAR, _C.USE_IMG, _C.USE_ACTION_MASK, _C.MAP_LEVEL
OBS_W, OBS_H, OBSTACLE_COLOR = _C.OBS_W, _C.OBS_H, _C.OBSTACLE_COLOR
_TYPE_LETTER = {0: "S", 1: "L", 2: "R"}


class Box(object):
    """the part of gym.spaces.Box the callers use: low/high/shape/dtype, sample(), seed()"""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        shape = np.asarray(low).shape if shape is None else tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), shape).copy()
        self.shape = tuple(shape)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class PATH(object):
    """info['path_to_dest'] as parking_agent.py:12-20 reads it (reeds_shepp.py:15-32)"""

    def __init__(self, lengths, ctypes, L):
        self.lengths, self.ctypes, self.L = lengths, ctypes, L
        self.x, self.y, self.yaw, self.directions = [], [], [], []


class _Map(object):
    """the fields of ParkingMapNormal / ParkingMapDLP the callers read (parking_map_normal.py:460-494)"""

    def __init__(self, level):
        self.kind = level  # what reset() generates: 'Normal' | 'Complex' | 'Extrem' (ParkingMapNormal) or 'dlp' (ParkingMapDLP)
        self.map_level, self.case_id = level, None
        self.start = self.dest = self.start_box = self.dest_box = None
        self.obstacles = []
        self.n_obstacle = 0
        self.xmin = self.xmax = self.ymin = self.ymax = 0

    def load(self, sc, i=0):
        self.start, self.dest = State(list(sc["start"][i])), State(list(sc["dest"][i]))
        self.start_box, self.dest_box = self.start.create_box(), self.dest.create_box()
        self.xmin, self.xmax, self.ymin, self.ymax = [float(v) for v in sc["bounds"][i]]
        self.obstacles = [Area(shape=Ring(sc["obs"][i, k, :nv]), subtype="obstacle", color=OBSTACLE_COLOR)
                          for k, nv in enumerate(sc["nverts"][i]) if nv]
        self.n_obstacle = len(self.obstacles)
        self.case_id = int(sc["case_id"][i]) if "case_id" in sc else None


class _GlobalNumpyRng(object):
    """numpy's global generator behind the three calls dlp.prepare_scene makes, in the reference's own order
    (parking_map_dlp.py:49, 62-64, 81-84), so `np.random.seed(s)` reproduces the reference's DLP resets"""

    @staticmethod
    def integers(lo, hi):
        return np.random.randint(lo, hi)

    @staticmethod
    def standard_normal(n):
        return np.array([np.random.randn() for _ in range(n)])

    @staticmethod
    def random():
        return np.random.random()


class CarParking(object):
    metadata = {"render_mode": ["human", "rgb_array"]}
    _OUT = ("lidar", "mask", "target", "reward", "done", "status", "reward_info", "pose", "substeps", "retreated", "rs_found", "rs_nseg",
            "rs_types", "rs_lengths", "rs_L")

    def __init__(self, render_mode=None, fps=FPS, verbose=True, use_lidar_observation=USE_LIDAR, use_img_observation=USE_IMG,
                 use_action_mask=USE_ACTION_MASK, device=0, seed=None):
        self.config = refconfig.validate(refconfig.load())
        self.verbose, self.fps = verbose, fps
        self.render_mode = "human" if render_mode is None else render_mode
        self.use_lidar_observation, self.use_img_observation, self.use_action_mask = use_lidar_observation, bool(use_img_observation), use_action_mask
        self.screen = self.matrix = self.clock = None
        self.is_open = True
        self.level = MAP_LEVEL
        self.t = 0.0
        self.k = None
        self.tgt_repr_size = 5
        self.vehicle = Vehicle(n_step=NUM_STEP, step_len=STEP_LENGTH)
        self.map = _Map(self.level)
        self.reward = self.prev_reward = self.accum_arrive_reward = 0.0
        self.action_space = Box(np.array([VALID_STEER[0], VALID_SPEED[0]]).astype(np.float32),
                                np.array([VALID_STEER[1], VALID_SPEED[1]]).astype(np.float32))
        self.observation_space = {}
        if use_action_mask:
            self.observation_space["action_mask"] = Box(0, 1, shape=(N_DISCRETE_ACTION,), dtype=np.float64)
        if use_img_observation:  # car_parking_base.py:93-98: (OBS_W // 4, OBS_H // 4, 3) uint8
            self.observation_space["img"] = Box(0, 255, shape=(OBS_W // 4, OBS_H // 4, 3), dtype=np.uint8)
            self.raw_img_shape = (OBS_W, OBS_H, 3)
        if use_lidar_observation:
            self.observation_space["lidar"] = Box(np.zeros(LIDAR_NUM), np.ones(LIDAR_NUM) * LIDAR_RANGE, shape=(LIDAR_NUM,), dtype=np.float64)
        self.observation_space["target"] = Box(np.array([0, -1, -1, -1, -1]), np.array([MAX_DIST_TO_DEST, 1, 1, 1, 1]), shape=(5,), dtype=np.float64)
        self._device = device
        self._seed = None if seed is None else int(seed)  # None: every reset draws its scene seed from numpy's global generator
        self._episode = 0
        self._backends = {}      # obstacle capacity (16 | 128) -> BatchedParkingEnv
        self._backend = None
        self._dlp_cases = None
        self._pending_scene = None
        self._last_obs = None

    # ---- scene handling ---------------------------------------------------------------------------
    DLP_PATH = "../data/dlp.data"  # ParkingMapDLP.default['path'] (parking_map_dlp.py:15-17), relative to src/

    def set_level(self, level=None):
        """car_parking_base.py:117-125 (None -> a default-level ParkingMapNormal, the level attribute unchanged)"""
        if level is None:
            self.map = _Map(MAP_LEVEL)
            return
        self.level = level
        self.map = _Map(self.level)

    def load_scene(self, scene):
        """Use this scene (dict with start/dest/bounds/obs/nverts, leading axis 1) on the next reset()."""
        self._pending_scene = {k: np.asarray(v) for k, v in scene.items()}

    def _scene_seed(self):
        self._episode += 1
        if self._seed is not None:
            return self._seed + self._episode
        return int(np.random.randint(0, 2 ** 31 - 1))

    def _next_scene(self, case_id, data_dir):
        if self._pending_scene is not None:
            sc, self._pending_scene = self._pending_scene, None
            return sc
        level = self.map.kind
        if level == "dlp":
            # ParkingMapDLP.reset (parking_map_dlp.py:38-86) through the shapely-free reader, numpy's global generator
            from hope_b200 import dlp
            if data_dir is not None:
                self._dlp_cases = dlp.read_dlp(data_dir)
            if self._dlp_cases is None:
                self._dlp_cases = dlp.read_dlp(self.DLP_PATH)
            n = len(self._dlp_cases)
            rng = _GlobalNumpyRng if self._seed is None else np.random.default_rng(self._scene_seed())
            cid = int(rng.integers(0, n)) if case_id is None else (int(case_id) % n if int(case_id) >= n else int(case_id))
            row = dlp.prepare_scene(self._dlp_cases[cid], rng)
            sc = {k: np.asarray(v)[None] for k, v in row.items()}
            sc["case_id"] = np.array([cid], dtype=np.int32)
            return sc
        sc = None
        for _ in range(64):  # hope_generate_scenes draws bay/parallel itself; honour an explicit case_id (parking_map_normal.py:475-480)
            sc = generate_scenes(1, level, self._scene_seed(), nthreads=1)
            if case_id not in (0, 1) or int(sc["case_id"][0]) == case_id or level == "Extrem":
                break
        return sc

    # ---- gym surface ------------------------------------------------------------------------------
    def reset(self, case_id=None, data_dir=None, level=None):
        self.reward = self.prev_reward = self.accum_arrive_reward = 0.0
        if level is not None:
            self.set_level(level)
        sc = self._next_scene(case_id, data_dir)
        self.map.load(sc)
        if self.map.kind == "dlp":  # ParkingMapDLP.reset labels the case (parking_map_dlp.py:86)
            from env.map_level import get_map_level
            self.map.map_level = get_map_level(self.map.start, self.map.dest, self.map.obstacles)
        cap = int(np.asarray(sc["nverts"]).shape[1])
        if cap not in self._backends:
            self._backends[cap] = BatchedParkingEnv(1, scenes=sc, device=self._device, auto_reset=False,
                                                    use_img_observation=self.use_img_observation, config=self.config)
        else:
            self._backends[cap].set_scene_pool(sc)
        self._backend = self._backends[cap]
        self.t = 1.0
        out = self._backend.reset_host(outputs=self._outputs())
        self.vehicle.reset(self.map.start)
        self.matrix = self.coord_transform_matrix()
        return self._unpack(out, moved=False)[0]

    def coord_transform_matrix(self):
        """car_parking_base.py:140-147"""
        k = self.config.K
        bx = 0.5 * (self.config.WIN_W - k * (self.map.xmax + self.map.xmin))
        by = 0.5 * (self.config.WIN_H - k * (self.map.ymax + self.map.ymin))
        self.k = k
        return [k, 0, 0, k, bx, by]

    def step(self, action=None):
        """car_parking_base.py:235-299.  `action` = [steer rad, speed m/s] (physical units; the wrapper rescales policy outputs),
        or None for a step without motion (what reset() ends with)."""
        if action is None:
            out = self._backend.step_host(None, outputs=self._outputs())
            self.t += 1
            return self._unpack(out, moved=False)
        a = np.asarray(action, dtype=np.float64).reshape(1, 2)
        out = self._backend.step_host(a, outputs=self._outputs(), raw_action=True)
        self.t += 1
        return self._unpack(out)

    def _step_unit(self, unit_action):
        """policy-scale action in [-1,1]^2 (what CarParkingWrapper.step receives); env_wrapper.py:37-50 runs on the device"""
        out = self._backend.step_host(np.asarray(unit_action, dtype=np.float64).reshape(1, 2), outputs=self._outputs())
        self.t += 1
        return self._unpack(out)

    def _outputs(self):
        return self._OUT + (("img",) if self.use_img_observation else ())

    def _unpack(self, out, moved=True):
        obs = {"img": None, "lidar": None, "target": out["target"][0].copy(), "action_mask": None}
        if self.use_img_observation:  # (64, 64, 3) float64 like process_img; the wrapper transposes to (3, 64, 64)
            obs["img"] = out["img"][0].transpose(1, 2, 0) / 255.0
        if self.use_lidar_observation:
            obs["lidar"] = out["lidar"][0].copy()
        if self.use_action_mask:
            obs["action_mask"] = out["mask"][0].copy()
        pose = out["pose"][0]
        veh = self.vehicle
        veh.state = State([pose[0], pose[1], pose[2], 0, 0])
        veh.box = veh.state.create_box()
        if moved and int(out["substeps"][0]) - int(out["retreated"][0]) >= 1:
            veh.trajectory.append(veh.state)  # one state per env step survives the pruning of car_parking_base.py:273-275
        status = Status(int(out["status"][0]))
        keys = ("time_cost", "rs_dist_reward", "dist_reward", "angle_reward", "box_union_reward")
        reward_info = OrderedDict((k, float(v)) for k, v in zip(keys, out["reward_info"][0]))
        info = OrderedDict({"reward_info": reward_info, "path_to_dest": None})
        if out["rs_found"][0]:
            n = int(out["rs_nseg"][0])
            info["path_to_dest"] = PATH([float(v) for v in out["rs_lengths"][0][:n]],
                                        [_TYPE_LETTER[int(c)] for c in out["rs_types"][0][:n]], float(out["rs_L"][0]))
        self._shaped_reward = float(out["reward"][0])
        self._last_obs = obs
        return obs, reward_info, status, info

    def render(self, mode="human"):
        """car_parking_base.py:383-411 returns the observation of the current state; there is no window to draw into"""
        assert mode in self.metadata["render_mode"]
        return None if self._last_obs is None else {k: (v.copy() if v is not None else None) for k, v in self._last_obs.items()}

    def seed(self, seed=None):
        self._seed = None if seed is None else int(seed)
        return [seed]

    def close(self):
        for b in self._backends.values():
            b.close()
        self._backends, self._backend = {}, None
        self.is_open = False
