"""TEST INFRASTRUCTURE ONLY — Python face of the CPU oracle (oracle/c/parking_oracle.c).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs.  The product package hope_b200/ must never import this module.

Contents
  build()            compile oracle/c/parking_oracle.c -> oracle/_build/libparking_oracle.so
  mask_tables()      numpy restatement of ActionMask.__init__ (model/action_mask.py:9-163) and
                     LidarSimlator.__init__ (env/lidar_simulator.py:14-53): ray tables, the two
                     vehicle-boundary offset vectors, dist_star[1200,42,10], upsample weights
  OracleEnv          N independent scenes stepped by the C oracle (float64, OpenMP over scenes)

Pinned against tests/golden (recorded from the unmodified reference by oracle/make_golden.py);
the GEOS predicates under it are restated, so parity with real shapely is UNPINNED.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

from . import geom

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libparking_oracle.so")
LIB_PATHS = {16: LIB_PATH, 128: os.path.join(HERE, "_build", "libparking_oracle_obs128.so")}

MAX_OBS, MAX_V, N_RAY, N_UP, N_ACT, N_ITER = 16, 4, 120, 1200, 42, 10
WHEEL_BASE, LIDAR_RANGE = 2.8, 10.0
BOX = [(-0.93, -1.94 / 2), (0.96 + 2.8, -1.94 / 2), (0.96 + 2.8, 1.94 / 2), (-0.93, 1.94 / 2)]  # configs.py:20-24
MAXC = math.tan(0.75) / WHEEL_BASE  # car_parking_base.py:422


def build(force=False):
    src = os.path.join(HERE, "c", "parking_oracle.c")
    if force or any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in LIB_PATHS.values()):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


def discrete_actions():
    """configs.py:108-115 — 21 steers from np.arange (the 11th is 4.44e-16, not 0) x {+1,-1}."""
    steers = np.arange(0.75, -(0.75 + 0.75 / 10), -0.75 / 10)
    return np.array([[s, 1.0] for s in steers] + [[s, -1.0] for s in steers])


def _boundary_offsets(cos_fn, sin_fn):
    """Distance from the rear-axle origin to where beam i leaves VehicleBox
    (LineString.intersection(ring).distance(origin): lidar_simulator.py:48-53, action_mask.py:21-29)."""
    ring = BOX + [BOX[0]]
    out = np.zeros(N_RAY)
    for i in range(N_RAY):
        end = (float(cos_fn(i * math.pi / N_RAY * 2) * LIDAR_RANGE), float(sin_fn(i * math.pi / N_RAY * 2) * LIDAR_RANGE))
        pts = geom.segment_ring_intersection_points((0.0, 0.0), end, ring)
        out[i] = min(math.hypot(px, py) for px, py in pts)
    return out


def upsample(x, rate=10):
    """action_mask.py:145-163 — circular linear interpolation along axis 0."""
    n = x.shape[0]
    j = np.arange(n * rate)
    ext = np.concatenate([x, x[:1]], axis=0)
    shp = (n * rate,) + (1,) * (x.ndim - 1)
    w_hi = ((j % rate) / rate).reshape(shp)
    w_lo = (1 - (j % rate) / rate).reshape(shp)
    return ext[j // rate] * w_lo + ext[j // rate + 1] * w_hi


def swept_boxes():
    """action_mask.py:84-112 -> corners[42, 10, 4, 2] of the box after k+1 arc steps of 0.05 m."""
    acts = discrete_actions()
    radius = 1 / (np.tan(acts[:, 0]) / WHEEL_BASE)
    cx = np.array([p[0] for p in BOX]).reshape(1, 4)
    cy = np.array([p[1] for p in BOX]).reshape(1, 4)
    centre_x = 0 - radius * np.sin(0)
    centre_y = 0 + radius * np.cos(0)
    dphi = 0.5 * acts[:, 1] / 10 / radius
    psi = 0
    frames = []
    for _ in range(N_ITER):
        psi = psi + dphi
        px = centre_x + radius * np.sin(psi)
        py = centre_y - radius * np.cos(psi)
        c = np.cos(psi).reshape(-1, 1)
        s = np.sin(psi).reshape(-1, 1)
        X = c * cx - s * cy + px.reshape(-1, 1)
        Y = s * cx + c * cy + py.reshape(-1, 1)
        frames.append(np.stack([X, Y], axis=-1))
    return np.stack(frames, axis=1)


def dist_star_table():
    """action_mask.py:114-143 with _intersect :31-82 -> dist_star[1200, 42, 10]."""
    boxes = swept_boxes()  # (42,10,4,2)
    idx = np.arange(N_RAY)
    ex = (np.cos(idx / N_RAY * 2 * np.pi) * (LIDAR_RANGE * 10)).reshape(-1, 1)
    ey = (np.sin(idx / N_RAY * 2 * np.pi) * (LIDAR_RANGE * 10)).reshape(-1, 1)
    zero = np.zeros_like(ex)
    # ray as line a x + b y + c = 0 through (0,0)->(ex,ey)
    a = ey - zero
    b = zero - ex
    c = zero * ex - zero * ey
    nxt = np.roll(boxes, -1, axis=2)  # edge = (vertex i+1) -> (vertex i)
    x1 = nxt[..., 0].reshape(1, -1); y1 = nxt[..., 1].reshape(1, -1)
    x2 = boxes[..., 0].reshape(1, -1); y2 = boxes[..., 1].reshape(1, -1)
    d = y2 - y1
    e = x1 - x2
    f = y1 * x2 - x1 * y2
    det = a * e - b * d
    par = det == 0
    det[par] = 1
    with np.errstate(all="ignore"):
        rx = (b * f - c * e) / det
        ry = (c * d - a * f) / det
    tol = 1e-8
    rx[rx > np.maximum(zero, ex) + tol] = np.inf
    rx[rx < np.minimum(zero, ex) - tol] = np.inf
    ry[ry > np.maximum(zero, ey) + tol] = np.inf
    ry[ry < np.minimum(zero, ey) - tol] = np.inf
    rx[rx > np.maximum(x1, x2) + tol] = np.inf
    rx[rx < np.minimum(x1, x2) - tol] = np.inf
    ry[ry > np.maximum(y1, y2) + tol] = np.inf
    ry[ry < np.minimum(y1, y2) - tol] = np.inf
    rx[par] = np.inf
    with np.errstate(all="ignore"):
        r = np.sqrt(rx * rx + ry * ry).reshape(N_RAY, N_ACT, N_ITER, 4)
    r[r == np.inf] = 0
    return upsample(r.max(axis=-1))


_TABLES = None


def mask_tables():
    global _TABLES
    if _TABLES is None:
        theta = np.array([i * math.pi / N_RAY * 2 for i in range(N_RAY)])
        r = np.arange(10)
        _TABLES = dict(
            ray_a=np.sin(theta), ray_b=-np.cos(theta),
            lidar_base=_boundary_offsets(math.cos, math.sin),
            mask_base=_boundary_offsets(np.cos, np.sin),
            dist_star=np.ascontiguousarray(dist_star_table()),
            w_lo=1 - (r % 10) / 10, w_hi=(r % 10) / 10, maxc=MAXC)
    return _TABLES


class _Tables(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in ("ray_a", "ray_b", "lidar_base", "mask_base", "dist_star", "w_lo", "w_hi")] + [("maxc", C.c_double)]


_IO_FIELDS = [
    ("start", C.c_double), ("dest", C.c_double), ("bounds", C.c_double), ("obs", C.c_double), ("nverts", C.c_int),
    ("pose", C.c_double), ("accum", C.c_double), ("t", C.c_int),
    ("lidar", C.c_double), ("mask", C.c_double), ("target", C.c_double), ("reward", C.c_double), ("reward_info", C.c_double),
    ("mask_steps", C.c_int), ("status", C.c_int), ("substeps", C.c_int), ("retreated", C.c_int),
    ("rs_found", C.c_int), ("rs_nseg", C.c_int), ("rs_ncand", C.c_int), ("rs_ntried", C.c_int), ("rs_T_last", C.c_int), ("rs_err", C.c_int),
    ("rs_types", C.c_uint8), ("rs_len", C.c_double), ("rs_L", C.c_double)]


class _IO(C.Structure):
    _fields_ = [(k, C.POINTER(t)) for k, t in _IO_FIELDS]


_LIBS = {}


def lib(max_obs=16):
    if max_obs not in _LIBS:
        build()
        _LIB = C.CDLL(LIB_PATHS[max_obs])
        assert _LIB.orc_max_obs() == max_obs
        _LIB.orc_step.restype = C.c_int
        _LIB.orc_step.argtypes = [C.c_int, C.POINTER(_IO), C.c_void_p, C.c_void_p, C.POINTER(_Tables), C.c_int, C.c_int]
        _LIB.orc_rs_all_paths.restype = C.c_int
        _LIB.orc_orient.restype = C.c_int
        _LIB.orc_orient.argtypes = [C.c_double] * 6
        _LIB.orc_seg_hit.restype = C.c_int
        _LIB.orc_clip_area.restype = C.c_double
        assert _LIB.orc_sizeof_io() == C.sizeof(_IO)
        _LIBS[max_obs] = _LIB
    return _LIBS[max_obs]


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleEnv(object):
    """N scenes, float64.  Scene arrays follow the SoA capacity of SURVEY §8: 16 obstacles x 4 vertices.

    step(actions) mirrors CarParkingWrapper.step (env_wrapper.py:73-81) per scene; reset_step()
    mirrors CarParking.reset's trailing self.step() (car_parking_base.py:127-138)."""

    OUT_SHAPES = dict(lidar=(N_RAY,), mask=(N_ACT,), target=(5,), reward=(), reward_info=(5,), mask_steps=(N_ACT,),
                      status=(), substeps=(), retreated=(), rs_found=(), rs_nseg=(), rs_ncand=(), rs_ntried=(),
                      rs_T_last=(), rs_err=(), rs_types=(5,), rs_len=(5,), rs_L=())

    def __init__(self, start, dest, bounds, obs, nverts, tables=None, nthreads=0):
        self.n = n = start.shape[0]
        MAX_OBS = int(np.asarray(nverts).shape[1])  # 16, or 128 for Dragon Lake Parking scenes
        f8 = lambda a, s: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape((n,) + s))
        self.start, self.dest, self.bounds = f8(start, (3,)), f8(dest, (3,)), f8(bounds, (4,))
        self.obs = f8(obs, (MAX_OBS, MAX_V, 2))
        self.nverts = np.ascontiguousarray(np.asarray(nverts, dtype=np.int32).reshape(n, MAX_OBS))
        self.pose = self.start.copy()
        self.accum = np.zeros(n)
        self.t = np.zeros(n, dtype=np.int32)
        self.nthreads = nthreads
        tb = tables or mask_tables()
        self._tb_arrays = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in tb.items() if k != "maxc"}
        self._tb = _Tables(*[_ptr(self._tb_arrays[k], C.c_double) for k in ("ray_a", "ray_b", "lidar_base", "mask_base", "dist_star", "w_lo", "w_hi")], tb["maxc"])
        types = dict(_IO_FIELDS)
        self.out = {}
        for k, s in self.OUT_SHAPES.items():
            dt = {C.c_double: np.float64, C.c_int: np.int32, C.c_uint8: np.uint8}[types[k]]
            self.out[k] = np.zeros((n,) + s, dtype=dt)
        arrs = dict(start=self.start, dest=self.dest, bounds=self.bounds, obs=self.obs, nverts=self.nverts,
                    pose=self.pose, accum=self.accum, t=self.t, **self.out)
        self._io = _IO(*[_ptr(arrs[k], t) for k, t in _IO_FIELDS])
        self._lib = lib(MAX_OBS)

    def reset_state(self, idx=None):
        idx = slice(None) if idx is None else idx
        self.pose[idx] = self.start[idx]
        self.accum[idx] = 0.0
        self.t[idx] = 0

    def step(self, actions=None, has_action=None, stages=3):
        a = None
        if actions is not None:
            a = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.n, 2)
        ha = None
        if has_action is not None:
            ha = np.ascontiguousarray(has_action, dtype=np.uint8)
        self._lib.orc_step(self.n, C.byref(self._io), a.ctypes.data if a is not None else None,
                           ha.ctypes.data if ha is not None else None, C.byref(self._tb), stages, self.nthreads)
        return self.out

    def reset_step(self, stages=1):
        self.reset_state()
        return self.step(None, stages=stages)


def rs_all_paths(q0, q1, maxc=MAXC, cap=48):
    L = lib()
    q0 = np.ascontiguousarray(q0, dtype=np.float64); q1 = np.ascontiguousarray(q1, dtype=np.float64)
    nseg = np.zeros(cap, dtype=np.int32); types = np.zeros((cap, 5), dtype=np.uint8); lens = np.zeros((cap, 5))
    Ls = np.zeros(cap); T = np.zeros(cap, dtype=np.int32); csum = np.zeros((cap, 3))
    head = np.zeros((cap, 3, 3)); tail = np.zeros((cap, 3, 3))
    n = L.orc_rs_all_paths(_ptr(q0, C.c_double), _ptr(q1, C.c_double), C.c_double(maxc), _ptr(nseg, C.c_int),
                           _ptr(types, C.c_uint8), _ptr(lens, C.c_double), _ptr(Ls, C.c_double), _ptr(T, C.c_int),
                           _ptr(csum, C.c_double), _ptr(head, C.c_double), _ptr(tail, C.c_double))
    err = n < 0
    if err:
        n = -n - 1
    return dict(n=n, err=err, nseg=nseg[:n], types=types[:n], lengths=lens[:n], L=Ls[:n], T=T[:n], csum=csum[:n],
                head=head[:n], tail=tail[:n])
