"""Small planar-geometry helpers for the host side of the drop-in facade (plain Python floats, no shapely).

The reference's scene classifier `get_map_level` (src/env/map_level.py:27-112) and its helpers call a handful of shapely
operations on a dozen rings per scene: point-to-ring and ring-to-ring distance, `equals`, the minimum rotated rectangle
of a point set, and polygon-vs-ring `intersects`.  These are their standard definitions; sizes are tiny (a label per
episode), so nothing here is on the hot path.  Rings are sequences of (x, y); a closing duplicate vertex is accepted and
ignored.
"""
import math


def open_ring(pts):
    pts = [(float(p[0]), float(p[1])) for p in pts]
    if len(pts) > 1 and pts[0] == pts[-1]:
        pts = pts[:-1]
    return pts


def _edges(ring):
    n = len(ring)
    return [(ring[i], ring[(i + 1) % n]) for i in range(n)]


def point_segment_distance(p, a, b):
    dx, dy = b[0] - a[0], b[1] - a[1]
    l2 = dx * dx + dy * dy
    if l2 == 0.0:
        return math.hypot(p[0] - a[0], p[1] - a[1])
    t = ((p[0] - a[0]) * dx + (p[1] - a[1]) * dy) / l2
    if t <= 0.0:
        return math.hypot(p[0] - a[0], p[1] - a[1])
    if t >= 1.0:
        return math.hypot(p[0] - b[0], p[1] - b[1])
    return abs((a[1] - p[1]) * dx - (a[0] - p[0]) * dy) / math.sqrt(l2)


def point_ring_distance(p, ring):
    """shapely `Point.distance(LinearRing)`: distance to the curve (a point inside the ring is NOT at distance 0)"""
    return min(point_segment_distance(p, a, b) for a, b in _edges(open_ring(ring)))


def _cross(o, a, b):
    return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])


def _on_segment(p, a, b):
    return min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1])


def segments_intersect(p1, p2, q1, q2):
    d1, d2 = _cross(p1, p2, q1), _cross(p1, p2, q2)
    d3, d4 = _cross(q1, q2, p1), _cross(q1, q2, p2)
    if ((d1 > 0 and d2 < 0) or (d1 < 0 and d2 > 0)) and ((d3 > 0 and d4 < 0) or (d3 < 0 and d4 > 0)):
        return True
    return ((d1 == 0 and _on_segment(q1, p1, p2)) or (d2 == 0 and _on_segment(q2, p1, p2)) or
            (d3 == 0 and _on_segment(p1, q1, q2)) or (d4 == 0 and _on_segment(p2, q1, q2)))


def rings_cross(ra, rb):
    ea, eb = _edges(open_ring(ra)), _edges(open_ring(rb))
    return any(segments_intersect(a0, a1, b0, b1) for a0, a1 in ea for b0, b1 in eb)


def ring_ring_distance(ra, rb):
    """shapely `LinearRing.distance(LinearRing)`: 0 when the curves share a point, else the closest vertex-to-edge gap"""
    ra, rb = open_ring(ra), open_ring(rb)
    if rings_cross(ra, rb):
        return 0.0
    ea, eb = _edges(ra), _edges(rb)
    best = math.inf
    for p in ra:
        best = min(best, min(point_segment_distance(p, a, b) for a, b in eb))
    for p in rb:
        best = min(best, min(point_segment_distance(p, a, b) for a, b in ea))
    return best


def rings_equal(ra, rb):
    """shapely `equals` for two simple rings: the same closed curve (any start vertex, either direction)"""
    ra, rb = open_ring(ra), open_ring(rb)
    if len(ra) != len(rb):
        return False
    n = len(ra)
    for seq in (rb, rb[::-1]):
        for s in range(n):
            if all(ra[i] == seq[(s + i) % n] for i in range(n)):
                return True
    return False


def convex_hull(pts):
    """Andrew's monotone chain; counter-clockwise, collinear points dropped"""
    pts = sorted(set((float(p[0]), float(p[1])) for p in pts))
    if len(pts) <= 2:
        return pts
    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and _cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and _cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return lower[:-1] + upper[:-1]


def minimum_rotated_rectangle(pts):
    """shapely 1.x `minimum_rotated_rectangle` (geometry/base.py): over the edges of the convex hull, the axis-parallel
    bounding rectangle in the edge's frame with the smallest area, transformed back.  Returns 4 corners."""
    hull = convex_hull(pts)
    if len(hull) < 3:
        return hull
    best = None
    for (ax, ay), (bx, by) in _edges(hull):
        dx, dy = bx - ax, by - ay
        length = math.sqrt(dx * dx + dy * dy)
        ux, uy = dx / length, dy / length
        vx, vy = -uy, ux
        us = [ux * x + uy * y for x, y in hull]
        vs = [vx * x + vy * y for x, y in hull]
        u0, u1, v0, v1 = min(us), max(us), min(vs), max(vs)
        area = (u1 - u0) * (v1 - v0)
        if best is None or area < best[0]:
            best = (area, [(ux * u + vx * v, uy * u + vy * v) for u, v in ((u0, v0), (u1, v0), (u1, v1), (u0, v1))])
    return best[1]


def point_in_convex(p, poly):
    """p inside or on the boundary of the convex polygon `poly` (either winding)"""
    sign = 0
    for a, b in _edges(poly):
        c = _cross(a, b, p)
        if c != 0:
            if sign == 0:
                sign = 1 if c > 0 else -1
            elif (c > 0) != (sign > 0):
                return False
    return True


def convex_polygon_intersects_ring(poly, ring):
    """shapely `Polygon.intersects(LinearRing)` for a convex polygon: the filled polygon and the curve share a point"""
    poly, ring = open_ring(poly), open_ring(ring)
    if any(point_in_convex(p, poly) for p in ring):
        return True
    return rings_cross(poly, ring)
