"""affine_transform as in shapely 1.x (`shapely/affinity.py`, 2-D branch): per coordinate
xp = a*x + b*y + xoff ; yp = d*x + e*y + yoff, evaluated left to right in float64."""
from shapely.geometry import Point, LineString, LinearRing


def affine_transform(geom, matrix):
    a, b, d, e, xoff, yoff = matrix
    out = []
    for x, y in geom.coords:
        out.append((a * x + b * y + xoff, d * x + e * y + yoff))
    if isinstance(geom, Point):
        return Point(out[0])
    cls = LinearRing if isinstance(geom, LinearRing) else LineString
    return cls(out)
