"""k_render stages the static screen rows of one image half through a fixed shared-memory buffer (render.cuh SWIN_BYTES) and
clamps silently if a window were larger, so the bound is checked here with the kernel's own integer camera map
(k_render_camera: the heading as a C float of degrees, isin / icos = (int)(sin, cos * 65536), car_parking_base.py:333-350 via
pygame's rotate()) over a dense sweep of headings and the worst alignment of the window against pixels and 64-pixel chunks."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _constant(name):
    src = open(os.path.join(ROOT, "hope_b200", "csrc", "render.cuh")).read()
    return int(re.search(r"constexpr int %s = (\d+);" % name, src).group(1))


def _rotation(h):
    ang = (h * (180.0 / np.pi)).astype(np.float32).astype(np.float64)
    quarter = np.fmod(ang, 90.0) == 0.0
    rad = ang * .01745329251994329
    isin = np.trunc(np.sin(rad) * 65536).astype(np.int64)
    icos = np.trunc(np.cos(rad) * 65536).astype(np.int64)
    q = (ang[quarter] / 90).astype(np.int64) % 4
    icos[quarter] = np.array([65536, 0, -65536, 0])[q]
    isin[quarter] = np.array([0, 65536, 0, -65536])[q]  # (signs do not matter for the extents)
    return isin, icos


def test_half_image_window_fits_the_staging_buffer():
    swin = _constant("SWIN_BYTES")
    h = np.concatenate([np.linspace(-np.pi, np.pi, 2_000_001), np.arange(-8, 9) * (np.pi / 2)])
    isin, icos = _rotation(h)
    # half of the sample lattice: crop pixels u = 1 .. 254, v = 1 .. 126 (or 129 .. 254): extents in 16.16 fixed point
    rx = np.abs(icos) * 253 + np.abs(isin) * 125
    ry = np.abs(isin) * 253 + np.abs(icos) * 125
    rows = ry // 65536 + 2                 # floor(a + r) - floor(a) + 1 <= floor(r) + 2
    width = rx // 65536 + 2
    chunks = (width - 1) // 64 + 2         # a run of w pixels straddles at most floor((w - 1) / 64) + 2 chunks of 64
    size = rows * chunks * 16
    assert size.max() <= swin, (int(size.max()), swin)
    assert chunks.max() <= 6 and rows.max() <= 285


def test_dynamic_windows_and_lattice_index_share_their_buffer():
    dwords, lat = _constant("DWORDS"), 2 * _constant("IMG")
    assert dwords * 4 >= lat * lat          # didx (lattice mode) aliases dwin
    assert dwords % (4 * _constant("THREADS")) == 0  # cleared as uint4 by every thread
