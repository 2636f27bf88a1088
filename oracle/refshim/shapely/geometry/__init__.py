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
        self.pts = _as_pairs([p.coords[0] if isinstance(p, Point) else p for p in pts])

    @property
    def minimum_rotated_rectangle(self):
        """shapely 1.x geometry/base.py: over the edges of the convex hull, the axis-parallel bounding rectangle in the
        edge's frame with the smallest area, transformed back (map_level.py:71, :98)"""
        return Polygon._from_open(_g.min_area_rectangle(self.pts))

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

    def __init__(self, coords=None):
        if coords is None:  # unpickling: shapely 1.x reduces a geometry to (class, (), wkb) and restores it in __setstate__
            self._coords = []
            return
        if isinstance(coords, LineString):
            coords = coords.coords
        pts = _as_pairs(coords)
        if self.closed and pts[0] != pts[-1]:
            pts.append(pts[0])
        self._coords = pts

    def __setstate__(self, state):
        """data/dlp.data pickles shapely-1.x LinearRings: the state is the WKB of a (closed) LineString"""
        if isinstance(state, dict):  # a stand-in object pickled by this module (oracle/make_dlp_golden.py)
            self.__dict__.update(state)
            return
        import struct
        b = bytes(state)
        order = "<" if b[0] == 1 else ">"
        gtype, n = struct.unpack_from(order + "II", b, 1)
        assert gtype & 0xFF == 2 and not gtype & 0x80000000, "2-D LineString WKB expected"
        flat = struct.unpack_from(order + "%dd" % (2 * n), b, 9)
        self._coords = [(flat[2 * i], flat[2 * i + 1]) for i in range(n)]

    def equals(self, other):
        return _g.rings_equal(self._coords, other._coords)

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

    @property
    def exterior(self):
        return LinearRing(self._open)

    def intersects(self, other):
        """filled convex polygon vs a curve (map_level.py:79, :106): a curve vertex inside or on the polygon, or crossing boundaries"""
        ring = self._open + [self._open[0]]
        if any(_g.point_in_convex(p, self._open) for p in other._coords):
            return True
        return _g.rings_intersect(ring, other._coords)
