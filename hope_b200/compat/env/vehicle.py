"""Mirror of the names the reference's callers import from env/vehicle.py (vehicle.py:13-18, configs.py:32-38)."""
from enum import Enum

VALID_SPEED = [-2.5, 2.5]
VALID_STEER = [-0.75, 0.75]
NUM_STEP = 10
STEP_LENGTH = 5e-2
WHEEL_BASE = 2.8


class Status(Enum):
    CONTINUE = 1
    ARRIVED = 2
    COLLIDED = 3
    OUTBOUND = 4
    OUTTIME = 5


class _Loc(object):
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)


class State(object):
    """pose holder with the attribute surface callers read: .loc.x/.loc.y, .heading, .get_pos()"""

    def __init__(self, raw_state):
        self.loc = _Loc(raw_state[0], raw_state[1])
        self.heading = float(raw_state[2])
        self.speed = float(raw_state[3]) if len(raw_state) > 3 else 0.0
        self.steering = float(raw_state[4]) if len(raw_state) > 4 else 0.0

    def get_pos(self):
        return (self.loc.x, self.loc.y, self.heading)


class KSModel(object):
    def __init__(self, wheel_base=WHEEL_BASE, step_len=STEP_LENGTH, n_step=NUM_STEP, speed_range=VALID_SPEED, angle_range=VALID_STEER):
        self.wheel_base, self.step_len, self.n_step = wheel_base, step_len, n_step
        self.speed_range, self.angle_range, self.mini_iter = speed_range, angle_range, 20


class Vehicle(object):
    """state mirror only: the motion itself happens on the GPU"""

    def __init__(self):
        self.kinetic_model = KSModel()
        self.state = None
        self.initial_state = None
        self.trajectory = []
