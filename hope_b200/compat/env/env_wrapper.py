"""CarParkingWrapper facade (env_wrapper.py:58-85) plus the module-level functions its constructor defaults to.

With the default functions (the only ones the reference's scripts use) action rescaling and reward shaping run on the
device (k_advance follows env_wrapper.py:10-50) and the wrapper only re-assembles the 4-tuple.  A caller-supplied
`action_func`, `reward_func` or `observation_func` is honoured the way the reference does it, on the host: the action
function's result goes to `CarParking.step` in physical units, the reward function sees `(obs, reward_info, status, info)`.
"""
import numpy as np

from hope_b200 import refconfig
from env.car_parking_base import CarParking
from env.vehicle import Status

_C = refconfig.load()
REWARD_WEIGHT, REWARD_RATIO = _C.REWARD_WEIGHT, _C.REWARD_RATIO


def reward_shaping(*args):
    """env_wrapper.py:10-35"""
    obs, reward_info, status, info = args
    if status == Status.CONTINUE:
        reward = 0
        for reward_type in REWARD_WEIGHT.keys():
            reward += REWARD_WEIGHT[reward_type] * reward_info[reward_type]
    elif status == Status.OUTTIME:
        reward = -1
    elif status == Status.ARRIVED:
        reward = 50
    else:  # OUTBOUND, COLLIDED
        reward = -50
    reward *= REWARD_RATIO
    info["status"] = status
    return obs, reward, status, info


def action_rescale(action, action_space, raw_action_range=(-1, 1), explore=True, epsilon=0.0):
    """env_wrapper.py:37-50"""
    action = np.clip(action, *raw_action_range)
    action = action * (action_space.high - action_space.low) / 2 + (action_space.high + action_space.low) / 2
    if explore and np.random.random() < epsilon:
        action = action_space.sample()
    return action


def observation_rescale(obs):
    """env_wrapper.py:52-55"""
    if obs["img"] is not None:
        obs["img"] = obs["img"].transpose((2, 0, 1))
    return obs


class CarParkingWrapper(object):
    def __init__(self, env: CarParking, action_func=action_rescale, reward_func=reward_shaping, observation_func=observation_rescale):
        self.env = env
        self.reward_func, self.action_func, self.obs_func = reward_func, action_func, observation_func
        self.observation_shape = {k: self.env.observation_space[k].shape for k in self.env.observation_space}
        if "img" in self.observation_shape:  # env_wrapper.py:68-71
            w, h, c = self.observation_shape["img"]
            self.observation_shape["img"] = (c, w, h)

    def __getattr__(self, name):  # gym.Wrapper forwards everything else to the wrapped env
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env

    def step(self, action=None):
        if action is None:  # env_wrapper.py:74-75
            return self.obs_func(self.env.step()[0])
        if self.action_func is action_rescale:  # clip + rescale on the device, bit-identical to the host expression
            obs, reward_info, status, info = self.env._step_unit(np.asarray(action, dtype=np.float64))
        else:
            obs, reward_info, status, info = self.env.step(self.action_func(action, self.env.action_space))
        if self.reward_func is reward_shaping:
            reward = self.env._shaped_reward  # k_advance evaluated env_wrapper.py:10-35
            info["status"] = status
        else:
            obs, reward, status, info = self.reward_func(obs, reward_info, status, info)
        obs = self.obs_func(obs)
        done = False if status == Status.CONTINUE else True
        return obs, reward, done, info

    def reset(self, *args):
        return self.obs_func(self.env.reset(*args))
