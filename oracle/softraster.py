"""TEST INFRASTRUCTURE ONLY — software restatement of the pygame subset behind the reference's image
observation (`car_parking_base.py:301-350`): Surface.fill / blit / subsurface / get_rect, draw.polygon
(filled and width=1), transform.rotate, image.tostring.

pygame is an unpinned, un-vendored dependency of the reference (`requirements.txt`) that is not installed
here and cannot be installed, so **parity with a real pygame build is unpinned**.  The algorithms below
restate the published pygame 2.x C sources (SDL2 software surfaces, 32-bit display format):

  draw.c   draw_fillpoly   integer scan conversion: vertices truncated to int, one pass per row y in
                           [miny, maxy]; an edge (y1 < y2 after ordering) contributes when y1 <= y < y2 or
                           when y == maxy == y2; its x is x1 + floor|ceil((y-y1)*(x2-x1) / (float)(y2-y1))
                           (floor for the 1st, 3rd.. crossing found, ceil for the 2nd, 4th..); crossings are
                           sorted and filled pairwise with inclusive horizontal runs; finally every horizontal
                           edge strictly between miny and maxy is drawn as a run.
  draw.c   draw_line       Bresenham with err = (dx > dy ? dx : -dy) / 2, both end points set; horizontal,
                           vertical and single-point lines special-cased (same pixels).
  draw.c   polygon(width=1) = lines(closed=True): consecutive segments, then last -> first.
  transform.c surf_rotate  angle parsed as a C float; multiples of 90 go through rotate90; otherwise the
                           destination is (int)max|+-cx+-sy| x (int)max|+-sx+-cy| and every destination pixel
                           takes the source pixel at 16.16 fixed-point coordinates (see rotate() below), or
                           the background = the source's top-left pixel when it falls outside.
  rect.c   Rect center setter: x = cx - w / 2, y = cy - h / 2 (C integer division).

Everything is exact integer / IEEE arithmetic, so the CUDA rasteriser (`k_render`) can be compared with
this bit for bit.  Pixels are RGB uint8 (alpha is never read on this path).
"""
import math

import numpy as np


def _c_int(v):
    """(int)double of C: truncation toward zero."""
    return int(v)


class Rect(object):
    def __init__(self, x, y, w, h):
        self.x, self.y, self.w, self.h = int(x), int(y), int(w), int(h)

    @property
    def center(self):
        return (self.x + self.w // 2, self.y + self.h // 2)

    @center.setter
    def center(self, c):
        self.x = int(c[0]) - (self.w >> 1)
        self.y = int(c[1]) - (self.h >> 1)

    @property
    def topleft(self):
        return (self.x, self.y)


class Surface(object):
    def __init__(self, size=(0, 0), _arr=None):
        if _arr is not None:
            self.arr = _arr
        else:
            w, h = int(size[0]), int(size[1])
            self.arr = np.zeros((h, w, 3), dtype=np.uint8)  # new SDL surfaces start black

    @property
    def w(self):
        return self.arr.shape[1]

    @property
    def h(self):
        return self.arr.shape[0]

    def get_size(self):
        return (self.w, self.h)

    def get_rect(self, **kw):
        r = Rect(0, 0, self.w, self.h)
        for k, v in kw.items():
            setattr(r, k, v)
        return r

    def fill(self, color):
        self.arr[:, :] = np.asarray(color[:3], dtype=np.uint8)

    def blit(self, src, dest):
        if isinstance(dest, Rect):
            dx, dy = dest.x, dest.y
        else:
            dx, dy = int(dest[0]), int(dest[1])
        x0, y0 = max(dx, 0), max(dy, 0)
        x1, y1 = min(dx + src.w, self.w), min(dy + src.h, self.h)
        if x1 <= x0 or y1 <= y0:
            return
        self.arr[y0:y1, x0:x1] = src.arr[y0 - dy:y1 - dy, x0 - dx:x1 - dx]

    def subsurface(self, *args):
        if len(args) == 2:
            (x, y), (w, h) = args
        else:
            (x, y), (w, h) = args[0]
        x, y, w, h = int(x), int(y), int(w), int(h)
        return Surface(_arr=self.arr[y:y + h, x:x + w])


# --------------------------------------------------------------------------------------------- draw.c
def _hline(surf, color, x1, y, x2):
    """drawhorzlineclip: inclusive run, clipped to the surface."""
    if y < 0 or y >= surf.h:
        return
    if x2 < x1:
        x1, x2 = x2, x1
    x1 = max(x1, 0)
    x2 = min(x2, surf.w - 1)
    if x2 < x1:
        return
    surf.arr[y, x1:x2 + 1] = color


def _vline(surf, color, x, y1, y2):
    if x < 0 or x >= surf.w:
        return
    if y2 < y1:
        y1, y2 = y2, y1
    y1 = max(y1, 0)
    y2 = min(y2, surf.h - 1)
    if y2 < y1:
        return
    surf.arr[y1:y2 + 1, x] = color


def _pixel(surf, color, x, y):
    if 0 <= x < surf.w and 0 <= y < surf.h:
        surf.arr[y, x] = color


def fillpoly_row_runs(px, py, y, miny, maxy):
    """The sorted crossing list draw_fillpoly builds for row y (pairs are filled inclusively)."""
    n = len(px)
    xs = []
    for i in range(n):
        ip = i - 1 if i else n - 1
        y1, y2 = py[ip], py[i]
        if y1 < y2:
            x1, x2 = px[ip], px[i]
        elif y1 > y2:
            y2, y1 = py[ip], py[i]
            x2, x1 = px[ip], px[i]
        else:
            continue
        if (y1 <= y < y2) or (y == maxy and y2 == maxy):
            q = np.float32((y - y1) * (x2 - x1)) / np.float32(y2 - y1)  # int product -> float, float division
            q = math.floor(q) if len(xs) % 2 == 0 else math.ceil(q)
            xs.append(int(q) + x1)
    xs.sort()
    return xs


def draw_fillpoly(surf, color, px, py):
    n = len(px)
    miny, maxy = min(py), max(py)
    if miny == maxy:
        _hline(surf, color, min(px), miny, max(px))
        return
    for y in range(miny, maxy + 1):
        xs = fillpoly_row_runs(px, py, y, miny, maxy)
        for i in range(0, len(xs) - 1, 2):
            _hline(surf, color, xs[i], y, xs[i + 1])
    for i in range(n):
        ip = i - 1 if i else n - 1
        y = py[i]
        if miny < y and py[ip] == y and y < maxy:
            _hline(surf, color, px[i], y, px[ip])


def line_pixels(x1, y1, x2, y2):
    """Pixels draw_line sets, in order (the special cases set the same pixels as the general loop would not:
    they are whole runs), as a list of (x, y)."""
    if x1 == x2 and y1 == y2:
        return [(x1, y1)]
    if y1 == y2:
        lo, hi = min(x1, x2), max(x1, x2)
        return [(x, y1) for x in range(lo, hi + 1)]
    if x1 == x2:
        lo, hi = min(y1, y2), max(y1, y2)
        return [(x1, y) for y in range(lo, hi + 1)]
    dx, sx = abs(x2 - x1), (1 if x1 < x2 else -1)
    dy, sy = abs(y2 - y1), (1 if y1 < y2 else -1)
    err = int((dx if dx > dy else -dy) / 2)  # C integer division truncates toward zero
    out = []
    while x1 != x2 or y1 != y2:
        out.append((x1, y1))
        e2 = err
        if e2 > -dx:
            err -= dy
            x1 += sx
        if e2 < dy:
            err += dx
            y1 += sy
    out.append((x2, y2))
    return out


def draw_line(surf, color, x1, y1, x2, y2):
    for x, y in line_pixels(x1, y1, x2, y2):
        _pixel(surf, color, x, y)


def polygon(surf, color, points, width=0):
    col = np.asarray(color[:3], dtype=np.uint8)
    px = [_c_int(p[0]) for p in points]
    py = [_c_int(p[1]) for p in points]
    if width == 0:
        draw_fillpoly(surf, col, px, py)
        return
    if width != 1:
        raise NotImplementedError("only width 0 and 1 are on the reference's path")
    n = len(px)
    for i in range(1, n):
        draw_line(surf, col, px[i - 1], py[i - 1], px[i], py[i])
    if n > 2:
        draw_line(surf, col, px[n - 1], py[n - 1], px[0], py[0])


# ---------------------------------------------------------------------------------------- transform.c
def rotate_params(w, h, angle_deg):
    """Fixed-point set-up of transform.c rotate() for a w x h source: dict with the destination size and the
    16.16 increments, or {'quarter': k} when the angle is a multiple of 90 (rotate90 path)."""
    angle = float(np.float32(angle_deg))  # "f" format: the angle is a C float
    if math.fmod(angle, 90.0) == 0.0:
        return {"quarter": int(angle) // 90 % 4}  # (int)angle, then normalised into [0, 360)
    rad = angle * .01745329251994329
    sa, ca = math.sin(rad), math.cos(rad)
    cx, cy, sx, sy = ca * w, ca * h, sa * w, sa * h
    nx = int(max(abs(cx + sy), abs(cx - sy), abs(-cx + sy), abs(-cx - sy)))
    ny = int(max(abs(sx + cy), abs(sx - cy), abs(-sx + cy), abs(-sx - cy)))
    return {
        "dw": nx, "dh": ny, "cy": ny // 2,
        "xd": (w - nx) << 15, "yd": (h - ny) << 15,
        "isin": int(sa * 65536), "icos": int(ca * 65536),
        "ax": (nx << 15) - int(ca * ((nx - 1) << 15)),
        "ay": (ny << 15) - int(sa * ((nx - 1) << 15)),
        "xmax": (w << 16) - 1, "ymax": (h << 16) - 1,
    }


def rotate(surf, angle_deg):
    p = rotate_params(surf.w, surf.h, angle_deg)
    if "quarter" in p:
        return Surface(_arr=np.ascontiguousarray(np.rot90(surf.arr, k=p["quarter"])))
    bg = surf.arr[0, 0].copy()
    ys = np.arange(p["dh"], dtype=np.int64)[:, None]
    xs = np.arange(p["dw"], dtype=np.int64)[None, :]
    dx = p["ax"] + p["isin"] * (p["cy"] - ys) + p["xd"] + p["icos"] * xs
    dy = p["ay"] - p["icos"] * (p["cy"] - ys) + p["yd"] + p["isin"] * xs
    outside = (dx < 0) | (dy < 0) | (dx > p["xmax"]) | (dy > p["ymax"])
    sxp = np.clip(dx >> 16, 0, surf.w - 1)
    syp = np.clip(dy >> 16, 0, surf.h - 1)
    out = surf.arr[syp, sxp]
    out[outside] = bg
    return Surface(_arr=out)


def tostring(surf, fmt="RGB"):
    assert fmt == "RGB"
    return np.ascontiguousarray(surf.arr).tobytes()
