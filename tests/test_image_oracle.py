"""CPU tests of the image-observation oracle: oracle/image_oracle.py replays the images
recorded from the unmodified reference (tests/golden/images_*.npz, made by oracle/make_golden.py --only images
on the raster restatement oracle/softraster.py + the real cv2) bit for bit, and its integer 4x down-sampling
equals the installed cv2.resize."""
import os

import numpy as np
import pytest

from oracle import image_oracle as io
from oracle import softraster as sr

LEVELS = ("Normal", "Complex", "Extrem")


@pytest.mark.parametrize("level", LEVELS)
def test_golden_images_replay_bit_exact(golden_dir, level):
    g = dict(np.load(os.path.join(golden_dir, f"images_{level}.npz")))
    n_ep = len(g["scene_start"])
    book = io.TrajectoryBook(n_ep)
    checked = stalled = 0
    last_ep = -1
    for k in range(len(g["ep"])):
        ep = int(g["ep"][k])
        rings = io.scene_rings(g["scene_obs"][ep], g["scene_nverts"][ep])
        if ep != last_ep:
            book.reset(ep, g["scene_start"][ep])
            img0 = io.render_observation(g["scene_start"][ep], g["scene_dest"][ep], g["scene_bounds"][ep], rings, book.traj[ep])
            assert np.array_equal(img0, g["scene_reset_img"][ep]), f"reset image of episode {ep}"
            last_ep = ep
        before = len(book.traj[ep])
        book.step(ep, g["pose"][k], g["substeps"][k], g["retreated"][k])
        stalled += len(book.traj[ep]) == before
        assert len(book.traj[ep]) == int(g["traj_len"][k]), "Vehicle.trajectory bookkeeping"
        if k % 3 == 0 or len(book.traj[ep]) == before:  # every third step (each render is ~30 ms of Python)
            img = io.render_observation(g["scene_start"][ep], g["scene_dest"][ep], g["scene_bounds"][ep], rings, book.traj[ep])
            assert np.array_equal(img, g["img"][k]), f"step {k}"
            checked += 1
    assert checked >= 50
    assert g["traj_len"].max() > io.TRAJ_RENDER_LEN  # the 20-box window is exercised


@pytest.mark.parametrize("tag", ["len7", "off"])
def test_golden_images_under_other_trajectory_settings(golden_dir, tag):
    """configs.py:86 TRAJ_RENDER_LEN = 7 and :105 RENDER_TRAJ = False, recorded from the unmodified reference
    (oracle/make_golden.py --only trajcfg): pins the oracle's traj_render_len parameter, against which the GPU tests
    check hope_set_render_traj."""
    g = dict(np.load(os.path.join(golden_dir, f"images_traj_{tag}.npz")))
    length = int(g["traj_render_len"])
    assert length == {"len7": 7, "off": 0}[tag]
    n_ep = len(g["scene_start"])
    book = io.TrajectoryBook(n_ep)
    last_ep, checked, differs = -1, 0, 0
    for k in range(len(g["ep"])):
        ep = int(g["ep"][k])
        rings = io.scene_rings(g["scene_obs"][ep], g["scene_nverts"][ep])
        scene = (g["scene_start"][ep], g["scene_dest"][ep], g["scene_bounds"][ep], rings)
        if ep != last_ep:
            book.reset(ep, g["scene_start"][ep])
            assert np.array_equal(io.render_observation(*scene, book.traj[ep], length), g["scene_reset_img"][ep])
            last_ep = ep
        book.step(ep, g["pose"][k], g["substeps"][k], g["retreated"][k])
        assert len(book.traj[ep]) == int(g["traj_len"][k])
        if k % 2 == 0:
            assert np.array_equal(io.render_observation(*scene, book.traj[ep], length), g["img"][k]), f"step {k}"
            checked += 1
            if k % 10 == 0 and len(book.traj[ep]) > 8:  # the setting matters: the default trail paints other bytes
                differs += not np.array_equal(io.render_observation(*scene, book.traj[ep]), g["img"][k])
    assert checked >= 25 and g["traj_len"].max() > 20 and differs > 0


def test_downsample_equals_cv2_resize():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for trial in range(6):
        if trial < 3:
            img = rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8)
        else:  # few flat colours, like the rendered scenes
            pal = rng.integers(0, 256, size=(6, 3), dtype=np.uint8)
            img = pal[rng.integers(0, 6, size=(256, 256)) // (1 + trial % 2)]
        assert np.array_equal(io.downsample4(img), cv2.resize(img, (64, 64)))


def test_fillpoly_covers_an_axis_aligned_box_inclusively():
    s = sr.Surface((20, 20))
    sr.polygon(s, (9, 9, 9), [(3.7, 4.2), (10.9, 4.9), (10.1, 8.3), (3.2, 8.8), (3.7, 4.2)])  # truncates to (3,4)-(10,8)
    want = np.zeros((20, 20), dtype=bool)
    want[4:9, 3:11] = True
    assert np.array_equal(s.arr[:, :, 0] == 9, want)


def test_outline_is_a_closed_one_pixel_loop():
    s = sr.Surface((40, 40))
    pts = [(5, 5), (30, 9), (26, 33), (2, 28), (5, 5)]
    sr.polygon(s, (1, 2, 3), pts, width=1)
    on = s.arr[:, :, 0] == 1
    for x, y in pts:
        assert on[y, x]
    # every set pixel has at least two set 8-neighbours or is a corner: no gaps along the loop
    ys, xs = np.nonzero(on)
    for x, y in zip(xs, ys):
        assert on[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2].sum() >= 3


def test_rotate_quarter_turns_and_general_angle():
    rng = np.random.default_rng(1)
    s = sr.Surface(_arr=rng.integers(0, 255, size=(50, 50, 3), dtype=np.uint8))
    assert np.array_equal(sr.rotate(s, 0.0).arr, s.arr)
    r = s
    for _ in range(4):
        r = sr.rotate(r, 90.0)
    assert np.array_equal(r.arr, s.arr)
    # a small general rotation keeps the centre pixel and grows the canvas
    g = sr.rotate(s, 10.0)
    assert g.w > 50 and g.h > 50
    assert np.array_equal(g.arr[g.h // 2, g.w // 2], s.arr[25, 25]) or np.array_equal(g.arr[g.h // 2, g.w // 2], s.arr[24, 24]) \
        or np.array_equal(g.arr[g.h // 2, g.w // 2], s.arr[24, 25]) or np.array_equal(g.arr[g.h // 2, g.w // 2], s.arr[25, 24])


def _bresenham_run_on_row(x1, y1, x2, y2, y):
    """The closed form k_render uses for the pixels a draw_line walk leaves on row y (render.cuh make_seg /
    paint_shape_row): x-major lines put pixel i on row offset ceil((i dy - err0) / dx), y-major lines put row offset j
    at column offset ceil((j dx + err0) / dy), with err0 = (dx > dy ? dx : -dy) / 2 truncated like C."""
    if y < min(y1, y2) or y > max(y1, y2):
        return None
    if y1 == y2:
        return (min(x1, x2), max(x1, x2))
    if x1 == x2:
        return (x1, x1)
    dx, dy = abs(x2 - x1), abs(y2 - y1)
    sx, sy = (1 if x1 < x2 else -1), (1 if y1 < y2 else -1)
    k = (y - y1) * sy
    if dx > dy:
        err0 = dx // 2
        ilo = max(((k - 1) * dx + err0) // dy + 1, 0)
        ihi = min((k * dx + err0) // dy, dx)
        a, b = x1 + sx * ilo, x1 + sx * ihi
        return (min(a, b), max(a, b))
    err0 = -(dy // 2)
    m = -((-(k * dx + err0)) // dy)  # ceil
    return (x1 + sx * m, x1 + sx * m)


def test_closed_form_bresenham_rows_equal_the_literal_walk():
    checked = 0
    for dxs in range(-33, 34):
        for dys in range(-33, 34):
            x1, y1 = 70, 70
            x2, y2 = x1 + dxs, y1 + dys
            rows = {}
            for x, y in sr.line_pixels(x1, y1, x2, y2):
                lo, hi = rows.get(y, (x, x))
                rows[y] = (min(lo, x), max(hi, x))
            for y in range(min(y1, y2) - 1, max(y1, y2) + 2):
                assert _bresenham_run_on_row(x1, y1, x2, y2, y) == rows.get(y), (dxs, dys, y)
                checked += 1
    assert checked > 80000
