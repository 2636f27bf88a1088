"""Mirror of the names the reference's callers import from env/vehicle.py (vehicle.py:13-157): Status, State, KSModel,
Vehicle and the constants `from configs import *` re-exports there (VALID_SPEED, NUM_STEP, ...).

The numbers come from the caller's own `configs` module when it is importable (hope_b200.refconfig), so editing the
reference's configs.py keeps taking effect; the motion itself happens on the GPU (k_advance), so KSModel / Vehicle only hold
the attributes the callers read (`vehicle.kinetic_model.step_len / n_step`, `vehicle.state.loc.x`, `vehicle.box`, ...).
"""
from enum import Enum
import math

from hope_b200 import refconfig

_C = refconfig.load()
VALID_SPEED, VALID_STEER = list(_C.VALID_SPEED), list(_C.VALID_STEER)
NUM_STEP, STEP_LENGTH, WHEEL_BASE = _C.NUM_STEP, _C.STEP_LENGTH, _C.WHEEL_BASE
FRONT_HANG, REAR_HANG, LENGTH, WIDTH = _C.FRONT_HANG, _C.REAR_HANG, _C.LENGTH, _C.WIDTH
COLOR_POOL = list(_C.COLOR_POOL)


class Status(Enum):
    CONTINUE = 1
    ARRIVED = 2
    COLLIDED = 3
    OUTBOUND = 4
    OUTTIME = 5


class Ring(object):
    """the part of shapely's LinearRing the callers touch: `.coords` (closed: first vertex repeated)"""

    def __init__(self, pts):
        pts = [(float(p[0]), float(p[1])) for p in pts]
        if len(pts) > 1 and pts[0] != pts[-1]:
            pts.append(pts[0])
        self.coords = pts

    @property
    def exterior(self):
        return self

    def distance(self, other):
        from hope_b200 import planar
        if isinstance(other, Loc):
            return planar.point_ring_distance((other.x, other.y), self.coords)
        return planar.ring_ring_distance(self.coords, other.coords)

    def intersects(self, other):
        from hope_b200 import planar
        return planar.rings_cross(self.coords, other.coords)

    def equals(self, other):
        from hope_b200 import planar
        return other is self or planar.rings_equal(self.coords, other.coords)


class Loc(object):
    """the part of shapely's Point the callers touch: .x, .y, .distance(), .coords"""
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)

    @property
    def coords(self):
        return [(self.x, self.y)]

    def distance(self, other):
        if isinstance(other, Loc):
            return math.hypot(self.x - other.x, self.y - other.y)
        return other.distance(self)


class State(object):
    """vehicle.py:21-39: pose holder; `.loc.x / .loc.y`, `.heading`, `.create_box()`, `.get_pos()`"""

    def __init__(self, raw_state):
        self.loc = Loc(raw_state[0], raw_state[1])
        self.heading = float(raw_state[2])
        self.speed = float(raw_state[3]) if len(raw_state) > 3 else 0.0
        self.steering = float(raw_state[4]) if len(raw_state) > 4 else 0.0

    def create_box(self):
        """corners rb, rf, lf, lb as `a*x + b*y + xoff` left to right (vehicle.py:32-36)"""
        c, s = math.cos(self.heading), math.sin(self.heading)
        ms = -s
        return Ring([(c * x + ms * y + self.loc.x, s * x + c * y + self.loc.y) for x, y in refconfig.load().VEHICLE_BOX])

    def get_pos(self):
        return (self.loc.x, self.loc.y, self.heading)


class KSModel(object):
    def __init__(self, wheel_base=WHEEL_BASE, step_len=STEP_LENGTH, n_step=NUM_STEP, speed_range=VALID_SPEED, angle_range=VALID_STEER):
        self.wheel_base, self.step_len, self.n_step = wheel_base, step_len, n_step
        self.speed_range, self.angle_range, self.mini_iter = speed_range, angle_range, 20


class Vehicle(object):
    """state mirror only (vehicle.py:99-157): the facade writes `.state`, `.box` and `.trajectory` after every step"""

    def __init__(self, wheel_base=WHEEL_BASE, step_len=STEP_LENGTH, n_step=NUM_STEP, speed_range=VALID_SPEED, angle_range=VALID_STEER):
        self.kinetic_model = KSModel(wheel_base, step_len, n_step, speed_range, angle_range)
        self.state = None
        self.initial_state = None
        self.box = None
        self.trajectory = []
        self.color = COLOR_POOL[0]
        self.v_max = self.v_min = None

    def reset(self, initial_state):
        self.initial_state = self.state = initial_state
        self.v_max = self.v_min = initial_state.speed
        self.box = initial_state.create_box()
        self.trajectory.clear()
        self.trajectory.append(initial_state)
