"""CarParking facade (car_parking_base.py:39-541 surface) over one env of the CUDA backend.

Same constructor, attributes and return values as the reference class for the lidar / target /
action-mask / image modalities.  Differences, all by construction of the backend:
  * scenes come from hope_generate_scenes (same distributions as parking_map_normal.py, own RNG),
    or from `load_scene()`; the global numpy RNG is not consumed;
  * no pygame window, no clock.tick (the reference sleeps to <= fps steps/s, :409);
  * the image (scope row f1) is rendered by k_render as uint8 and divided by 255.0 here, which is exactly
    what Obs_Processor.process_img returns (observation_processor.py:13-17); there is no pygame surface.
"""
from collections import OrderedDict

import numpy as np

from hope_b200 import capi
from hope_b200.batched_env import BatchedParkingEnv, generate_scenes
from env.vehicle import State, Status, Vehicle, VALID_SPEED, VALID_STEER

LIDAR_NUM, LIDAR_RANGE, N_DISCRETE_ACTION, MAX_DIST_TO_DEST = 120, 10.0, 42, 20
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


class PATH(object):
    """info['path_to_dest'] as parking_agent.py:12-20 reads it"""

    def __init__(self, lengths, ctypes, L):
        self.lengths, self.ctypes, self.L = lengths, ctypes, L
        self.x, self.y, self.yaw, self.directions = [], [], [], []


class _Area(object):
    def __init__(self, coords):
        self.shape = self
        self.coords = [tuple(p) for p in coords] + [tuple(coords[0])]
        self.subtype, self.color = "obstacle", (150, 150, 150, 255)

    def get_shape(self):
        return np.array(self.coords)


class _Map(object):
    def __init__(self, level):
        self.map_level, self.case_id = level, None
        self.start = self.dest = None
        self.obstacles = []
        self.xmin = self.xmax = self.ymin = self.ymax = 0

    def load(self, sc, i=0):
        self.start, self.dest = State(list(sc["start"][i]) + [0, 0]), State(list(sc["dest"][i]) + [0, 0])
        self.xmin, self.xmax, self.ymin, self.ymax = [float(v) for v in sc["bounds"][i]]
        self.obstacles = [_Area(sc["obs"][i, k, :nv]) for k, nv in enumerate(sc["nverts"][i]) if nv]
        self.n_obstacle = len(self.obstacles)
        self.case_id = int(sc["case_id"][i]) if "case_id" in sc else None


class CarParking(object):
    metadata = {"render_mode": ["human", "rgb_array"]}
    _OUT = ("lidar", "mask", "target", "reward", "done", "status", "reward_info", "pose", "rs_found", "rs_nseg",
            "rs_types", "rs_lengths", "rs_L")

    def __init__(self, render_mode=None, fps=100, verbose=True, use_lidar_observation=True, use_img_observation=True,
                 use_action_mask=True, device=0, seed=None):
        self.verbose, self.fps = verbose, fps
        self.render_mode = "human" if render_mode is None else render_mode
        self.use_lidar_observation, self.use_img_observation, self.use_action_mask = use_lidar_observation, bool(use_img_observation), use_action_mask
        self.level = "Normal"
        self.t = 0.0
        self.vehicle = Vehicle()
        self.map = _Map(self.level)
        self.action_space = Box(np.array([VALID_STEER[0], VALID_SPEED[0]]).astype(np.float32),
                                np.array([VALID_STEER[1], VALID_SPEED[1]]).astype(np.float32))
        self.observation_space = {}
        if use_action_mask:
            self.observation_space["action_mask"] = Box(0, 1, shape=(N_DISCRETE_ACTION,), dtype=np.float64)
        if use_img_observation:  # car_parking_base.py:93-98: (OBS_W // 4, OBS_H // 4, 3) uint8
            self.observation_space["img"] = Box(0, 255, shape=(64, 64, 3), dtype=np.uint8)
            self.raw_img_shape = (256, 256, 3)
        if use_lidar_observation:
            self.observation_space["lidar"] = Box(np.zeros(LIDAR_NUM), np.ones(LIDAR_NUM) * LIDAR_RANGE, shape=(LIDAR_NUM,), dtype=np.float64)
        self.observation_space["target"] = Box(np.array([0, -1, -1, -1, -1]), np.array([MAX_DIST_TO_DEST, 1, 1, 1, 1]), shape=(5,), dtype=np.float64)
        self._device = device
        self._seed = int(np.random.SeedSequence(seed).generate_state(1)[0]) if seed is None else int(seed)
        self._episode = 0
        self._backends = {}      # obstacle capacity (16 | 128) -> BatchedParkingEnv
        self._backend = None
        self._dlp_cases = None
        self._pending_scene = None

    # ---- scene handling ---------------------------------------------------------------------------
    DLP_PATH = "../data/dlp.data"  # ParkingMapDLP.default['path'] (parking_map_dlp.py:15-17), relative to src/

    def set_level(self, level=None):
        self.level = "Normal" if level is None else level
        self.map = _Map(self.level)

    def load_scene(self, scene):
        """Use this scene (dict with start/dest/bounds/obs/nverts, leading axis 1) on the next reset()."""
        self._pending_scene = {k: np.asarray(v) for k, v in scene.items()}

    def _next_scene(self, case_id):
        if self._pending_scene is not None:
            sc, self._pending_scene = self._pending_scene, None
            return sc
        if self.level == "dlp":  # ParkingMapDLP.reset (parking_map_dlp.py:38-86) through the shapely-free reader
            from hope_b200 import dlp
            if self._dlp_cases is None:
                self._dlp_cases = dlp.read_dlp(self.DLP_PATH)
            self._episode += 1
            rng = np.random.default_rng(self._seed + self._episode)
            cid = int(rng.integers(0, len(self._dlp_cases))) if case_id is None else int(case_id) % len(self._dlp_cases)
            sc = dlp.prepare_scenes(self._dlp_cases, [cid], seed=self._seed + self._episode)
            return sc
        for _ in range(64):  # hope_generate_scenes draws bay/parallel itself; honour an explicit case_id
            self._episode += 1
            sc = generate_scenes(1, self.level, self._seed + self._episode, nthreads=1)
            if case_id not in (0, 1) or int(sc["case_id"][0]) == case_id or self.level == "Extrem":
                return sc
        return sc

    # ---- gym surface ------------------------------------------------------------------------------
    def reset(self, case_id=None, data_dir=None, level=None):
        if level is not None:
            self.set_level(level)
        sc = self._next_scene(case_id)
        self.map.map_level = self.level
        self.map.load(sc)
        cap = int(np.asarray(sc["nverts"]).shape[1])
        if cap not in self._backends:
            self._backends[cap] = BatchedParkingEnv(1, scenes=sc, device=self._device, auto_reset=False,
                                                    use_img_observation=self.use_img_observation)
        else:
            self._backends[cap].set_scene_pool(sc)
        self._backend = self._backends[cap]
        self.t = 1.0
        out = self._backend.reset_host(outputs=self._outputs())
        self.vehicle.initial_state = self.map.start
        return self._unpack(out)[0]

    def step(self, action=None):
        if action is None:
            raise NotImplementedError("step() without an action is only used by reset() in the reference")
        a = np.asarray(action, dtype=np.float64).reshape(2)
        # the backend takes the policy-scale action and rescales on the device (env_wrapper.py:37-50);
        # CarParking.step receives physical units, so map back: steer/0.75, speed/2.5
        return self._step_unit(np.array([a[0] / VALID_STEER[1], a[1] / VALID_SPEED[1]]))

    def _step_unit(self, unit_action):
        """policy-scale action in [-1,1]^2 (what CarParkingWrapper.step receives), no round trip through physical units"""
        out = self._backend.step_host(np.asarray(unit_action, dtype=np.float64).reshape(1, 2), outputs=self._outputs())
        self.t += 1
        return self._unpack(out)

    def _outputs(self):
        return self._OUT + (("img",) if self.use_img_observation else ())

    def _unpack(self, out):
        obs = {"img": None, "lidar": None, "target": out["target"][0].copy(), "action_mask": None}
        if self.use_img_observation:  # (64, 64, 3) float64 like process_img; the wrapper transposes to (3, 64, 64)
            obs["img"] = out["img"][0].transpose(1, 2, 0) / 255.0
        if self.use_lidar_observation:
            obs["lidar"] = out["lidar"][0].copy()
        if self.use_action_mask:
            obs["action_mask"] = out["mask"][0].copy()
        pose = out["pose"][0]
        self.vehicle.state = State([pose[0], pose[1], pose[2], 0, 0])
        status = Status(int(out["status"][0]))
        keys = ("time_cost", "rs_dist_reward", "dist_reward", "angle_reward", "box_union_reward")
        reward_info = OrderedDict((k, float(v)) for k, v in zip(keys, out["reward_info"][0]))
        info = OrderedDict({"reward_info": reward_info, "path_to_dest": None})
        if out["rs_found"][0]:
            n = int(out["rs_nseg"][0])
            info["path_to_dest"] = PATH([float(v) for v in out["rs_lengths"][0][:n]],
                                        [_TYPE_LETTER[int(c)] for c in out["rs_types"][0][:n]], float(out["rs_L"][0]))
        self._shaped_reward = float(out["reward"][0])
        return obs, reward_info, status, info

    def render(self, mode="human"):
        return None

    def close(self):
        for b in self._backends.values():
            b.close()
        self._backends, self._backend = {}, None
