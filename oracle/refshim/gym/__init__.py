"""gym stand-in: only what car_parking_base.py / env_wrapper.py touch."""
from . import spaces, error


class Env(object):
    metadata = {}

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, 'unwrapped', self.env)
