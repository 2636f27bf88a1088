"""TEST INFRASTRUCTURE ONLY — planar-geometry restatement of the shapely/GEOS subset the
reference's ParkingEnv step calls.  Nothing in the product path (hope_b200/) may import this.

shapely is an un-vendored, unpinned third-party dependency of the reference
(requirements.txt:2; evidence for 1.x in SURVEY.md §8c) and is NOT installed here, so the
predicates below restate the published (OGC / GEOS) semantics of each call the hot path makes:

  LinearRing.intersects(LinearRing)     car_parking_base.py:153-158, parking_map_normal.py:159,185,189,217,222
  Polygon.intersection(Polygon).area    car_parking_base.py:164-170, 216-220
  Point.distance(Point)                 car_parking_base.py:207-211, 294
  ring.distance(Point | ring)           lidar_simulator.py:69, parking_map_normal.py:121,149
  LineString.intersection(ring).distance(Point)   lidar_simulator.py:51, action_mask.py:27
  affine_transform(geom, [a,b,d,e,xo,yo])         vehicle.py:36, lidar_simulator.py:68

PARITY UNPINNED against real GEOS: the reference ships no tests or golden vectors.  The
orientation predicate here is *exact* (float filter, then rational arithmetic), which is what
GEOS' robust DD orientation computes except for inputs far outside this domain; areas and
distances follow the textbook formulas and agree with GEOS to a few ulp, not bit-for-bit.
"""
from fractions import Fraction
import math

_EPS = 2.0 ** -53
_ORIENT_ERRBOUND = (3.0 + 16.0 * _EPS) * _EPS


def orient(ax, ay, bx, by, cx, cy):
    """Exact sign of the orientation determinant of (a, b, c): +1 ccw, -1 cw, 0 collinear."""
    detl = (ax - cx) * (by - cy)
    detr = (ay - cy) * (bx - cx)
    det = detl - detr
    bound = _ORIENT_ERRBOUND * (abs(detl) + abs(detr))
    if det > bound:
        return 1
    if det < -bound:
        return -1
    F = Fraction
    d = (F(ax) - F(cx)) * (F(by) - F(cy)) - (F(ay) - F(cy)) * (F(bx) - F(cx))
    return (d > 0) - (d < 0)


def _on_segment_box(px, py, ax, ay, bx, by):
    return min(ax, bx) <= px <= max(ax, bx) and min(ay, by) <= py <= max(ay, by)


def segments_intersect(p1, p2, q1, q2):
    """Closed segments p1p2 and q1q2 share at least one point (crossing, touch or overlap)."""
    if max(p1[0], p2[0]) < min(q1[0], q2[0]) or max(q1[0], q2[0]) < min(p1[0], p2[0]):
        return False
    if max(p1[1], p2[1]) < min(q1[1], q2[1]) or max(q1[1], q2[1]) < min(p1[1], p2[1]):
        return False
    o1 = orient(*p1, *p2, *q1)
    o2 = orient(*p1, *p2, *q2)
    o3 = orient(*q1, *q2, *p1)
    o4 = orient(*q1, *q2, *p2)
    if o1 * o2 < 0 and o3 * o4 < 0:
        return True
    if o1 == 0 and _on_segment_box(*q1, *p1, *p2):
        return True
    if o2 == 0 and _on_segment_box(*q2, *p1, *p2):
        return True
    if o3 == 0 and _on_segment_box(*p1, *q1, *q2):
        return True
    if o4 == 0 and _on_segment_box(*p2, *q1, *q2):
        return True
    return False


def rings_intersect(ring_a, ring_b):
    """ring_*: closed coordinate lists (first == last). Boundary-vs-boundary only."""
    for i in range(len(ring_a) - 1):
        for j in range(len(ring_b) - 1):
            if segments_intersect(ring_a[i], ring_a[i + 1], ring_b[j], ring_b[j + 1]):
                return True
    return False


def point_segment_distance(px, py, ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    l2 = dx * dx + dy * dy
    if l2 == 0.0:
        return math.hypot(px - ax, py - ay)
    r = ((px - ax) * dx + (py - ay) * dy) / l2
    if r <= 0.0:
        return math.hypot(px - ax, py - ay)
    if r >= 1.0:
        return math.hypot(px - bx, py - by)
    s = ((ay - py) * dx - (ax - px) * dy) / l2
    return abs(s) * math.sqrt(l2)


def point_ring_distance(px, py, ring):
    return min(point_segment_distance(px, py, *ring[i], *ring[i + 1]) for i in range(len(ring) - 1))


def ring_ring_distance(ring_a, ring_b):
    if rings_intersect(ring_a, ring_b):
        return 0.0
    best = math.inf
    for i in range(len(ring_a) - 1):
        for j in range(len(ring_b) - 1):
            a0, a1, b0, b1 = ring_a[i], ring_a[i + 1], ring_b[j], ring_b[j + 1]
            best = min(best,
                       point_segment_distance(*a0, *b0, *b1), point_segment_distance(*a1, *b0, *b1),
                       point_segment_distance(*b0, *a0, *a1), point_segment_distance(*b1, *a0, *a1))
    return best


def shoelace_area(poly):
    """Unsigned area of an open vertex list."""
    n = len(poly)
    if n < 3:
        return 0.0
    s = 0.0
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        s += x0 * y1 - x1 * y0
    return abs(s) * 0.5


def _signed_area(poly):
    s = 0.0
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        s += x0 * y1 - x1 * y0
    return 0.5 * s


def convex_clip(subject, clip):
    """Sutherland-Hodgman: open vertex list `subject` clipped to convex `clip` (any winding)."""
    if _signed_area(clip) < 0:
        clip = clip[::-1]
    out = list(subject)
    n = len(clip)
    for i in range(n):
        if not out:
            break
        ax, ay = clip[i]
        bx, by = clip[(i + 1) % n]
        ex, ey = bx - ax, by - ay
        inp, out = out, []
        m = len(inp)
        for k in range(m):
            px, py = inp[k]
            qx, qy = inp[(k + 1) % m]
            sp = ex * (py - ay) - ey * (px - ax)
            sq = ex * (qy - ay) - ey * (qx - ax)
            if sp >= 0:
                out.append((px, py))
                if sq < 0:
                    t = sp / (sp - sq)
                    out.append((px + t * (qx - px), py + t * (qy - py)))
            elif sq >= 0:
                t = sp / (sp - sq)
                out.append((px + t * (qx - px), py + t * (qy - py)))
    return out


def convex_intersection_area(poly_a, poly_b):
    return shoelace_area(convex_clip(poly_a, poly_b))


def segment_ring_intersection_points(p0, p1, ring):
    """Intersection points of segment p0p1 with each ring edge (transversal cases; a collinear
    overlap contributes the overlap's end points)."""
    pts = []
    for i in range(len(ring) - 1):
        q0, q1 = ring[i], ring[i + 1]
        if not segments_intersect(p0, p1, q0, q1):
            continue
        rx, ry = p1[0] - p0[0], p1[1] - p0[1]
        sx, sy = q1[0] - q0[0], q1[1] - q0[1]
        den = rx * sy - ry * sx
        if den != 0.0:
            t = ((q0[0] - p0[0]) * sy - (q0[1] - p0[1]) * sx) / den
            t = min(1.0, max(0.0, t))
            pts.append((p0[0] + t * rx, p0[1] + t * ry))
        else:
            for c in (q0, q1):
                if _on_segment_box(*c, *p0, *p1):
                    pts.append(c)
            for c in (p0, p1):
                if _on_segment_box(*c, *q0, *q1):
                    pts.append(c)
    return pts


def ring_centroid(ring):
    """Centroid of a lineal geometry (`LinearRing.centroid`, car_parking_base.py:328): GEOS
    `Centroid::addLineSegments` — length-weighted mean of the segment mid-points, accumulated in ring
    order; zero-length segments skipped; length = sqrt(dx*dx + dy*dy)."""
    total = 0.0
    sx = 0.0
    sy = 0.0
    for i in range(len(ring) - 1):
        (x0, y0), (x1, y1) = ring[i], ring[i + 1]
        dx, dy = x0 - x1, y0 - y1
        seg = math.sqrt(dx * dx + dy * dy)
        if seg == 0.0:
            continue
        total += seg
        sx += seg * ((x0 + x1) / 2)
        sy += seg * ((y0 + y1) / 2)
    return sx / total, sy / total


def rings_equal(ring_a, ring_b):
    """`equals` for two simple closed rings (first == last): the same curve from any start vertex, in either direction."""
    a, b = list(ring_a[:-1]), list(ring_b[:-1])
    if len(a) != len(b):
        return False
    n = len(a)
    for seq in (b, b[::-1]):
        for s in range(n):
            if all(a[i] == seq[(s + i) % n] for i in range(n)):
                return True
    return False


def convex_hull(pts):
    """counter-clockwise hull (gift wrapping from the lowest-leftmost point), collinear points dropped"""
    pts = list(dict.fromkeys((float(x), float(y)) for x, y in pts))
    if len(pts) < 3:
        return pts
    start = min(pts, key=lambda p: (p[1], p[0]))
    hull, cur = [], start
    while True:
        hull.append(cur)
        nxt = pts[0] if pts[0] != cur else pts[1]
        for p in pts:
            if p == cur:
                continue
            o = orient(cur[0], cur[1], nxt[0], nxt[1], p[0], p[1])
            if o < 0 or (o == 0 and math.hypot(p[0] - cur[0], p[1] - cur[1]) > math.hypot(nxt[0] - cur[0], nxt[1] - cur[1])):
                nxt = p
        cur = nxt
        if cur == start:
            return hull


def min_area_rectangle(pts):
    """shapely 1.x `minimum_rotated_rectangle`: min over hull edges of the bounding rectangle in the edge's frame (4 corners)"""
    hull = convex_hull(pts)
    best = None
    n = len(hull)
    for i in range(n):
        (ax, ay), (bx, by) = hull[i], hull[(i + 1) % n]
        length = math.sqrt((bx - ax) ** 2 + (by - ay) ** 2)
        ux, uy = (bx - ax) / length, (by - ay) / length
        vx, vy = -uy, ux
        us = [ux * x + uy * y for x, y in hull]
        vs = [vx * x + vy * y for x, y in hull]
        area = (max(us) - min(us)) * (max(vs) - min(vs))
        if best is None or area < best[0]:
            best = (area, [(ux * u + vx * v, uy * u + vy * v) for u, v in
                           ((min(us), min(vs)), (max(us), min(vs)), (max(us), max(vs)), (min(us), max(vs)))])
    return best[1]


def point_in_convex(p, poly):
    """p inside or on the convex polygon (open vertex list, either winding), exact orientation signs"""
    sign = 0
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        o = orient(a[0], a[1], b[0], b[1], p[0], p[1])
        if o != 0:
            if sign == 0:
                sign = o
            elif o != sign:
                return False
    return True
