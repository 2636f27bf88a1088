"""shapely.geometry stand-in: Point, LineString, LinearRing, Polygon, MultiPoint."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
import geom as _g  # oracle/geom.py

from .base import BaseGeometry


def _as_pairs(coords):
    return [(float(c[0]), float(c[1])) for c in coords]


class Point(BaseGeometry):
    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            if isinstance(a, Point):
                x, y = a.x, a.y
            else:
                x, y = a[0], a[1]
        else:
            x, y = args[0], args[1]
        self.x = float(x)
        self.y = float(y)

    @property
    def coords(self):
        return [(self.x, self.y)]

    def distance(self, other):
        if isinstance(other, Point):
            return math.hypot(self.x - other.x, self.y - other.y)
        return other.distance(self)


class MultiPoint(BaseGeometry):
    def __init__(self, pts=()):
        self.pts = _as_pairs(pts)

    @property
    def coords(self):
        return list(self.pts)

    @property
    def is_empty(self):
        return len(self.pts) == 0

    def distance(self, other):
        if isinstance(other, Point):
            return min(math.hypot(px - other.x, py - other.y) for px, py in self.pts)
        raise NotImplementedError


class LineString(BaseGeometry):
    closed = False

    def __init__(self, coords):
        if isinstance(coords, LineString):
            coords = coords.coords
        pts = _as_pairs(coords)
        if self.closed and pts[0] != pts[-1]:
            pts.append(pts[0])
        self._coords = pts

    @property
    def coords(self):
        return list(self._coords)

    def intersects(self, other):
        return _g.rings_intersect(self._coords, other._coords)

    @property
    def centroid(self):
        return Point(_g.ring_centroid(self._coords))

    def distance(self, other):
        if isinstance(other, Point):
            return _g.point_ring_distance(other.x, other.y, self._coords)
        return _g.ring_ring_distance(self._coords, other._coords)

    def intersection(self, other):
        pts = []
        for i in range(len(self._coords) - 1):
            pts += _g.segment_ring_intersection_points(self._coords[i], self._coords[i + 1], other._coords)
        return MultiPoint(pts)


class LinearRing(LineString):
    closed = True


class Polygon(BaseGeometry):
    def __init__(self, shell=None):
        if isinstance(shell, LineString):
            pts = shell.coords
        else:
            pts = _as_pairs(shell)
        if pts[0] == pts[-1]:
            pts = pts[:-1]
        self._open = pts

    @property
    def area(self):
        return _g.shoelace_area(self._open)

    def intersection(self, other):
        return Polygon._from_open(_g.convex_clip(self._open, other._open))

    @classmethod
    def _from_open(cls, pts):
        p = cls.__new__(cls)
        p._open = list(pts)
        return p
