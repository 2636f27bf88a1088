"""Host-side constant tables of the ParkingEnv step, uploaded once with hope_upload_tables.

Built with numpy in float64 exactly the way the reference's constructors do, so the thresholds
the mask sweep compares against carry the same bits:
  ray tables        env/lidar_simulator.py:86-88   theta_i = i*pi/120*2, A = sin, B = -cos
  own-box offsets   env/lidar_simulator.py:48-53 (math.cos/sin end points) and
                    model/action_mask.py:21-29 (np.cos/np.sin end points): distance from the
                    rear-axle origin to where beam i leaves VehicleBox
  dist_star         model/action_mask.py:84-163: for the 42 discrete actions (configs.py:108-115)
                    x 10 arc steps, farthest ray/box intersection, x10 circular upsample

Every constant comes from the configs snapshot of `refconfig.load()` (the caller's own `configs` module when importable,
SURVEY.md §5): WHEEL_BASE, the VehicleBox corners, LIDAR_RANGE, VALID_STEER, PRECISION / step_speed.
"""
import math

import numpy as np

from . import refconfig

N_RAY, N_ACT, N_ITER, UPSAMPLE = 120, 42, 10, 10  # compiled sizes (include/hope_b200.h); refconfig.validate checks configs against them


def discrete_actions(cfg=None):
    cfg = cfg or refconfig.load()
    return np.array(cfg.discrete_actions, dtype=np.float64)  # configs.py:108-115


def _exit_distance(ex, ey, box):
    """Where the segment (0,0)->(ex,ey) crosses the box ring, as a distance from the origin.
    The origin is strictly inside the box, so exactly one edge is crossed."""
    best = math.inf
    for k in range(4):
        qx, qy = box[k]
        sx, sy = box[(k + 1) % 4] - box[k]
        den = ex * sy - ey * sx
        if den == 0.0:
            continue
        t = (qx * sy - qy * sx) / den       # along the beam
        u = (qx * ey - qy * ex) / den       # along the edge
        if 0.0 <= t <= 1.0 and -1e-12 <= u <= 1.0 + 1e-12:
            t = min(1.0, max(0.0, t))
            best = min(best, math.hypot(0.0 + t * ex, 0.0 + t * ey))
    return best


def own_box_offsets(cos_fn, sin_fn, cfg=None):
    cfg = cfg or refconfig.load()
    out = np.zeros(N_RAY)
    for i in range(N_RAY):
        ang = i * math.pi / N_RAY * 2
        out[i] = _exit_distance(float(cos_fn(ang) * cfg.LIDAR_RANGE), float(sin_fn(ang) * cfg.LIDAR_RANGE), cfg.VEHICLE_BOX)
    return out


def circular_upsample(x, rate=UPSAMPLE):
    n = x.shape[0]
    j = np.arange(n * rate)
    wrap = np.concatenate([x, x[0:1]], axis=0)
    bshape = (n * rate,) + (1,) * (x.ndim - 1)
    lo = (1 - (j % rate) / rate).reshape(bshape)
    hi = ((j % rate) / rate).reshape(bshape)
    return wrap[j // rate, ...] * lo + wrap[j // rate + 1, ...] * hi


def swept_box_corners(cfg=None):
    cfg = cfg or refconfig.load()
    act = discrete_actions(cfg)
    radius = 1 / (np.tan(act[:, 0]) / cfg.WHEEL_BASE)
    bx = cfg.VEHICLE_BOX[:, 0].reshape(1, -1)
    by = cfg.VEHICLE_BOX[:, 1].reshape(1, -1)
    ox = 0 - radius * np.sin(0)
    oy = 0 + radius * np.cos(0)
    dpsi = 0.5 * act[:, 1] / 10 / radius
    psi = 0
    out = np.zeros((N_ACT, N_ITER, 4, 2))
    for k in range(N_ITER):
        psi = psi + dpsi
        px = ox + radius * np.sin(psi)
        py = oy - radius * np.cos(psi)
        c = np.cos(psi).reshape(-1, 1)
        s = np.sin(psi).reshape(-1, 1)
        out[:, k, :, 0] = c * bx - s * by + px.reshape(-1, 1)
        out[:, k, :, 1] = s * bx + c * by + py.reshape(-1, 1)
    return out


def dist_star(cfg=None):
    cfg = cfg or refconfig.load()
    corners = swept_box_corners(cfg)
    ray = np.arange(N_RAY)
    far = cfg.LIDAR_RANGE * 10
    x2r = (np.cos(ray / N_RAY * 2 * np.pi) * far).reshape(-1, 1)
    y2r = (np.sin(ray / N_RAY * 2 * np.pi) * far).reshape(-1, 1)
    x1r = np.zeros_like(x2r)
    y1r = np.zeros_like(y2r)
    a = y2r - y1r
    b = x1r - x2r
    c = y1r * x2r - x1r * y2r
    head = np.roll(corners, -1, axis=2)  # each box edge runs from the next corner back to this one
    x1 = head[..., 0].reshape(1, -1); y1 = head[..., 1].reshape(1, -1)
    x2 = corners[..., 0].reshape(1, -1); y2 = corners[..., 1].reshape(1, -1)
    d = y2 - y1
    e = x1 - x2
    f = y1 * x2 - x1 * y2
    det = a * e - b * d
    parallel = det == 0
    det[parallel] = 1
    with np.errstate(all="ignore"):
        ix = (b * f - c * e) / det
        iy = (c * d - a * f) / det
    tol = 1e-8
    inf = np.inf
    ix[ix > np.maximum(x1r, x2r) + tol] = inf
    ix[ix < np.minimum(x1r, x2r) - tol] = inf
    iy[iy > np.maximum(y1r, y2r) + tol] = inf
    iy[iy < np.minimum(y1r, y2r) - tol] = inf
    ix[ix > np.maximum(x1, x2) + tol] = inf
    ix[ix < np.minimum(x1, x2) - tol] = inf
    iy[iy > np.maximum(y1, y2) + tol] = inf
    iy[iy < np.minimum(y1, y2) - tol] = inf
    ix[parallel] = inf
    with np.errstate(all="ignore"):
        reach = np.sqrt(ix * ix + iy * iy)
    reach[reach == inf] = 0
    reach = reach.reshape(N_RAY, N_ACT, N_ITER, 4).max(axis=-1)
    return np.ascontiguousarray(circular_upsample(reach))


_CACHE = {}


def host_tables(cfg=None):
    """dict of contiguous float64 arrays in hope_upload_tables order, for the configs snapshot `cfg` (default: refconfig.load())."""
    cfg = refconfig.validate(cfg or refconfig.load())
    key = (float(cfg.WHEEL_BASE), cfg.VEHICLE_BOX.tobytes(), float(cfg.LIDAR_RANGE), float(cfg.VALID_STEER[-1]), int(cfg.PRECISION),
           float(cfg.step_speed))
    if key not in _CACHE:
        theta = np.array([i * math.pi / N_RAY * 2 for i in range(N_RAY)])
        r = np.arange(UPSAMPLE)
        _CACHE[key] = dict(
            ray_a=np.ascontiguousarray(np.sin(theta)), ray_b=np.ascontiguousarray(-np.cos(theta)),
            lidar_base=own_box_offsets(math.cos, math.sin, cfg), mask_base=own_box_offsets(np.cos, np.sin, cfg),
            dist_star=dist_star(cfg),
            w_lo=np.ascontiguousarray(1 - (r % UPSAMPLE) / UPSAMPLE), w_hi=np.ascontiguousarray((r % UPSAMPLE) / UPSAMPLE))
    return _CACHE[key]
