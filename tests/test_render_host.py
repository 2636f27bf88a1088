"""CPU fuzz test of k_render's scan conversion: the per-row helpers of hope_b200/csrc/render.cuh (the code every row
owner runs on the GPU) are compiled with g++ through tests/render_host_harness.cpp and compared, pixel for pixel,
with the raster oracle (oracle/softraster.py = restated pygame draw_fillpoly / draw_line) on thousands of shapes,
including the ones the GPU parity tests rarely or never visit: triangles, slivers, one-pixel-high shapes, horizontal
edges, concave and self-intersecting quads (two runs per row), shapes partly or wholly off screen, outlines."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import softraster as sr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WIN = 500


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("render_host") / "render_host.so")
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    subprocess.run([gxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", out,
                    os.path.join(ROOT, "tests", "render_host_harness.cpp")], check=True, env=env)
    lib = C.CDLL(out)
    lib.render_host_paint.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.render_host_covers.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def _oracle(pts, width):
    s = sr.Surface((WIN, WIN))
    sr.polygon(s, (7, 7, 7), [tuple(p) for p in pts] + [tuple(pts[0])], width)  # closed list, like shapely's coords
    return (s.arr[:, :, 0] == 7).astype(np.uint8)


def _paint(lib, pts, outline):
    px = np.ascontiguousarray([p[0] for p in pts], dtype=np.int32)
    py = np.ascontiguousarray([p[1] for p in pts], dtype=np.int32)
    img = np.zeros((WIN, WIN), dtype=np.uint8)
    assert lib.render_host_paint(px.ctypes.data, py.ctypes.data, len(pts), 1, 1 if outline else 0, img.ctypes.data) == 0
    return img


def _shapes(rng, n):
    out = []
    for i in range(n):
        kind = i % 8
        if kind in (0, 1):      # rotated rectangle, vehicle- to wall-sized, anywhere incl. partly off screen
            cx, cy = rng.uniform(-30, 530, size=2)
            w, h = rng.uniform(2, 120), rng.uniform(2, 60)
            th = rng.uniform(0, 2 * np.pi) if kind == 0 else rng.choice([0.0, np.pi / 2, np.pi, 1e-3])
            c, s_ = np.cos(th), np.sin(th)
            pts = [(cx + c * a - s_ * b, cy + s_ * a + c * b) for a, b in ((-w, -h), (w, -h), (w, h), (-w, h))]
        elif kind == 2:         # triangle
            pts = rng.uniform(-20, 520, size=(3, 2))
        elif kind == 3:         # any four points: concave or self-intersecting quads give two runs on a row
            c0 = rng.uniform(50, 450, size=2)
            pts = c0 + rng.uniform(-60, 60, size=(4, 2))
        elif kind == 4:         # sliver / one pixel high / repeated vertices
            x0, y0 = rng.integers(0, 480, size=2)
            pts = [(x0, y0), (x0 + rng.integers(0, 40), y0 + rng.integers(0, 2)), (x0 + rng.integers(0, 40), y0 + rng.integers(0, 2)), (x0, y0)]
        elif kind == 5:         # axis-aligned with horizontal edges in the middle (an L-like quad cannot exist; a trapezoid can)
            x0, y0 = rng.integers(20, 400, size=2)
            pts = [(x0, y0), (x0 + rng.integers(5, 60), y0), (x0 + rng.integers(5, 80), y0 + rng.integers(1, 50)), (x0 - rng.integers(0, 15), y0 + rng.integers(1, 50))]
        elif kind == 6:         # arrow head (concave)
            x0, y0 = rng.integers(60, 420, size=2)
            d = rng.integers(8, 50)
            pts = [(x0, y0 - d), (x0 + d, y0 + d), (x0, y0 + rng.integers(-d + 1, d)), (x0 - d, y0 + d)]
        else:                   # small shapes near the screen corner (the background probe's pixel)
            pts = rng.uniform(-6, 12, size=(4, 2))
        out.append([(int(p[0]), int(p[1])) for p in pts])   # pygame truncates; the harness takes the integers
    return out


def test_fill_rows_equal_draw_fillpoly(host):
    rng = np.random.default_rng(11)
    two_runs = 0
    for pts in _shapes(rng, 1600):
        want = _oracle(pts, 0)
        got = _paint(host, pts, False)
        assert np.array_equal(got, want), pts
        # the exact point query agrees with the painted pixels on screen
        cov = np.zeros((WIN, WIN), dtype=np.uint8)
        px = np.ascontiguousarray([p[0] for p in pts], dtype=np.int32); py = np.ascontiguousarray([p[1] for p in pts], dtype=np.int32)
        host.render_host_covers(px.ctypes.data, py.ctypes.data, len(pts), cov.ctypes.data)
        assert np.array_equal(cov, want), pts
        rows = want[(want.sum(axis=1) > 0)]
        two_runs += int(((np.diff(rows.astype(np.int8), axis=1) == 1).sum(axis=1) + rows[:, 0] > 1).any())
    assert two_runs > 20   # rows with two separate runs were exercised


def test_outline_rows_equal_draw_lines(host):
    rng = np.random.default_rng(12)
    shapes = [s for s in _shapes(rng, 900) if len(s) == 4]
    assert len(shapes) > 600
    for pts in shapes:
        assert np.array_equal(_paint(host, pts, True), _oracle(pts, 1)), pts
