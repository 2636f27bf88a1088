"""CPU fuzz test of the float64 helpers the kernels run (hope_b200/csrc/hope_device.cuh, compiled with g++ through
tests/device_host_harness.cpp) against the rational-arithmetic oracle oracle/geom.py, on adversarial inputs: exactly
collinear and touching configurations, perturbations of one ulp, large offsets that defeat a plain float64
determinant.  The collision booleans of the env step (car_parking_base.py:153-158) are built from these predicates."""
import ctypes as C
import math
import os
import shutil
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from oracle import geom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("device_host") / "device_host.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "tests", "host_stubs"),
                    "-o", out, os.path.join(ROOT, "tests", "device_host_harness.cpp")], check=True, env=env)
    lib = C.CDLL(out)
    lib.dev_orient.argtypes = [C.c_void_p, C.c_void_p]
    lib.dev_segments_touch.argtypes = [C.c_void_p, C.c_void_p]
    lib.dev_quad_clip_area.argtypes = [C.c_void_p] * 4
    lib.dev_quad_clip_area.restype = C.c_double
    lib.dev_rs_M.argtypes = [C.c_double]; lib.dev_rs_M.restype = C.c_double
    lib.dev_pi_2_pi.argtypes = [C.c_double]; lib.dev_pi_2_pi.restype = C.c_double
    return lib


def _ulp_nudge(rng, v):
    k = int(rng.integers(-2, 3))
    for _ in range(abs(k)):
        v = np.nextafter(v, np.inf if k > 0 else -np.inf)
    return float(v)


def _segment_cases(rng, n):
    """Pairs of segments, a third of them degenerate on purpose."""
    for i in range(n):
        off = rng.choice([0.0, 1.0, 37.25, 1e3, 1e6]) * rng.choice([-1.0, 1.0])
        p1 = rng.uniform(-10, 10, size=2) + off
        p2 = rng.uniform(-10, 10, size=2) + off
        kind = i % 6
        if kind == 0:      # general position
            q1, q2 = rng.uniform(-10, 10, size=2) + off, rng.uniform(-10, 10, size=2) + off
        elif kind == 1:    # q1 exactly on the carrier of p (rational parameter), q2 anywhere
            t = rng.choice([0.0, 0.25, 0.5, 1.0, 1.5, -0.5])
            q1 = p1 + t * (p2 - p1)
            q2 = rng.uniform(-10, 10, size=2) + off
        elif kind == 2:    # shared end point
            q1, q2 = p2.copy(), rng.uniform(-10, 10, size=2) + off
        elif kind == 3:    # collinear overlap / disjoint on the same carrier
            t1, t2 = rng.uniform(-1, 2, size=2)
            q1, q2 = p1 + t1 * (p2 - p1), p1 + t2 * (p2 - p1)
        elif kind == 4:    # axis-aligned (the parking scenes are full of these), touching or one ulp apart
            p1 = np.array([off + 1.0, off + 2.0]); p2 = np.array([off + 5.0, off + 2.0])
            x = off + rng.choice([1.0, 3.0, 5.0])
            q1 = np.array([x, off + 2.0]); q2 = np.array([x, off + 2.0 + rng.uniform(0.1, 3)])
        else:              # near miss: an end point one or two ulps off the other segment
            t = rng.uniform(0, 1)
            q1 = p1 + t * (p2 - p1)
            q2 = rng.uniform(-10, 10, size=2) + off
        if kind in (1, 2, 3, 4, 5) and rng.random() < 0.5:
            q1 = np.array([_ulp_nudge(rng, q1[0]), _ulp_nudge(rng, q1[1])])
        yield [float(v) for v in (*p1, *p2, *q1, *q2)]


def test_orientation_sign_is_exact(dev):
    rng = np.random.default_rng(21)
    fb = C.c_ulonglong(0)
    zeros = 0
    for s in _segment_cases(rng, 6000):
        a = np.array(s[:6], dtype=np.float64)
        want = geom.orient(*s[:6])
        d = (Fraction(s[0]) - Fraction(s[4])) * (Fraction(s[3]) - Fraction(s[5])) - (Fraction(s[1]) - Fraction(s[5])) * (Fraction(s[2]) - Fraction(s[4]))
        assert want == (d > 0) - (d < 0)
        assert dev.dev_orient(a.ctypes.data, C.byref(fb)) == want, s
        zeros += want == 0
    assert zeros > 200 and fb.value > 200   # exact zeros and the expansion fallback were both exercised


def test_segments_touch_equals_the_rational_oracle(dev):
    rng = np.random.default_rng(22)
    fb = C.c_ulonglong(0)
    hits = 0
    for s in _segment_cases(rng, 12000):
        a = np.array(s, dtype=np.float64)
        want = geom.segments_intersect(s[0:2], s[2:4], s[4:6], s[6:8])
        assert bool(dev.dev_segments_touch(a.ctypes.data, C.byref(fb))) == want, s
        hits += want
    assert 2000 < hits < 10000


def test_quad_clip_area_close_to_the_oracle(dev):
    rng = np.random.default_rng(23)
    for _ in range(3000):
        def box(cx, cy, w, h, th):
            c, s_ = math.cos(th), math.sin(th)
            return [(cx + c * a - s_ * b, cy + s_ * a + c * b) for a, b in ((-w, -h), (w, -h), (w, h), (-w, h))]  # counter-clockwise
        s = box(*rng.uniform(-3, 3, size=2), rng.uniform(0.5, 3), rng.uniform(0.5, 2), rng.uniform(0, 6.3))
        c = box(*rng.uniform(-3, 3, size=2), 2.345, 0.97, rng.uniform(0, 6.3))
        sx, sy = (np.array([p[k] for p in s]) for k in (0, 1))
        cx, cy = (np.array([p[k] for p in c]) for k in (0, 1))
        got = dev.dev_quad_clip_area(sx.ctypes.data, sy.ctypes.data, cx.ctypes.data, cy.ctypes.data)
        want = geom.convex_intersection_area(s, c)
        assert abs(got - want) <= 1e-12 * max(1.0, want), (s, c)


def test_angle_wraps_follow_python_semantics(dev):
    rng = np.random.default_rng(24)
    PI = math.pi
    for th in list(rng.uniform(-30, 30, size=4000)) + [0.0, PI, -PI, 2 * PI, -2 * PI, 3 * PI, -3 * PI, 1e-300, -1e-300]:
        th = float(th)
        phi = th % (2.0 * PI)            # reeds_shepp.py:581-592 (CPython floored modulo)
        if phi < -PI:
            phi += 2.0 * PI
        if phi > PI:
            phi -= 2.0 * PI
        assert dev.dev_rs_M(th) == phi, th
        w = th                           # reeds_shepp.py:561-568
        while w > PI:
            w -= 2.0 * PI
        while w < -PI:
            w += 2.0 * PI
        assert dev.dev_pi_2_pi(th) == w, th
