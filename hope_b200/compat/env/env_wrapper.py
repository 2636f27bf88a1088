"""CarParkingWrapper facade (env_wrapper.py:58-85).  Reward shaping already happened on the device
(hope_out.reward follows env_wrapper.py:10-35), so the wrapper only re-assembles the 4-tuple."""
import numpy as np

from env.car_parking_base import CarParking
from env.vehicle import Status


class CarParkingWrapper(object):
    def __init__(self, env: CarParking, action_func=None, reward_func=None, observation_func=None):
        if action_func is not None or reward_func is not None or observation_func is not None:
            raise NotImplementedError("custom action/reward/observation functions would have to run on the device")
        self.env = env
        self.observation_shape = {k: self.env.observation_space[k].shape for k in self.env.observation_space}
        if "img" in self.observation_shape:  # env_wrapper.py:68-71
            w, h, c = self.observation_shape["img"]
            self.observation_shape["img"] = (c, w, h)

    @staticmethod
    def _rescale(obs):  # observation_rescale, env_wrapper.py:52-55
        if obs["img"] is not None:
            obs["img"] = obs["img"].transpose((2, 0, 1))
        return obs

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action=None):
        if action is None:
            raise NotImplementedError
        # action_rescale (env_wrapper.py:37-50: clip, scale to steer/speed units) runs on the device
        obs, reward_info, status, info = self.env._step_unit(np.asarray(action, dtype=np.float64))
        info["status"] = status
        return self._rescale(obs), self.env._shaped_reward, status != Status.CONTINUE, info

    def reset(self, *args):
        return self._rescale(self.env.reset(*args))
