"""`get_map_level` — the reference's scene-difficulty classifier (src/env/map_level.py:27-112), without shapely.

Imported by the reference's own evaluation code (`evaluation/eval_utils.py:13`, used at :99) and by `ParkingMapDLP.reset`
(`parking_map_dlp.py:86`) to label a Dragon Lake Parking case "Normal" / "Complex" / "Extrem".  Same decision tree and the
same thresholds (`configs.py:43-76`, read through hope_b200.refconfig); the shapely calls it makes are replaced by the
helpers of hope_b200.planar:

    Point.distance(ring)                    -> point_ring_distance            (_get_nearest_obstacle, :157-171)
    ring.distance(ring)                     -> ring_ring_distance             (_has_enough_space, :173-199)
    ring.equals(ring)                       -> rings_equal                    (:77, :104, :161)
    MultiPoint(...).minimum_rotated_rectangle, Polygon.intersects(ring)
                                            -> minimum_rotated_rectangle, convex_polygon_intersects_ring   (:71-81, :98-107)

Accepts what the facade and the reference pass: a list of `Area` (anything with `.shape.coords`), of polygons (`.exterior`),
of rings (`.coords`), or of plain vertex arrays.
"""
import math

from hope_b200 import planar, refconfig

LEVEL_NORMAL, LEVEL_COMPLEX, LEVEL_EXTREM = "Normal", "Complex", "Extrem"
DEBUG = False


def _cfg():
    return refconfig.load()


def _ring_of(obj):
    if hasattr(obj, "shape") and hasattr(obj.shape, "coords"):
        return planar.open_ring(obj.shape.coords)          # Area
    if hasattr(obj, "exterior"):
        return planar.open_ring(obj.exterior.coords)       # Polygon
    if hasattr(obj, "coords"):
        return planar.open_ring(obj.coords)                # LinearRing
    return planar.open_ring(obj)                           # (n, 2) array


def _pose(state):
    if hasattr(state, "get_pos"):
        return tuple(float(v) for v in state.get_pos())
    return float(state[0]), float(state[1]), float(state[2])


def _box(pose):
    """State.create_box (vehicle.py:32-36): corners rb, rf, lf, lb"""
    x, y, h = pose
    c, s = math.cos(h), math.sin(h)
    ms = -s
    return [(c * bx + ms * by + x, s * bx + c * by + y) for bx, by in _cfg().VEHICLE_BOX]


def _midpoint(p, q):
    return ((p[0] + q[0]) / 2, (p[1] + q[1]) / 2)


def _translate(pt, heading, dist):
    return (pt[0] + math.cos(heading) * dist, pt[1] + math.sin(heading) * dist)


def _nearest_obstacle(pt, rings, max_min_dist, taken):
    """:157-171: index of the nearest ring closer than max_min_dist that was not already detected, else None"""
    best, best_k = max_min_dist, None
    for k, ring in enumerate(rings):
        if any(t is not None and (t == k or planar.rings_equal(ring, rings[t])) for t in taken):
            continue
        d = planar.point_ring_distance(pt, ring)
        if d < best:
            best, best_k = d, k
    return best_k


def _surrounding(pose, rings):
    """:14-25: nearest obstacle to the left, right, front and back edge midpoints of the box (each obstacle used once)"""
    rb, rf, lf, lb = _box(pose)
    pts = [_midpoint(lf, lb), _midpoint(rf, rb), _midpoint(lf, rf), _midpoint(lb, rb)]
    found = []
    for pt in pts:
        found.append(_nearest_obstacle(pt, rings, _cfg().LENGTH / 2, found))
    return found  # left, right, front, back


def _has_enough_space(pose, rings, width=None, length=None):
    """:173-199"""
    c = _cfg()
    box = _box(pose)
    ok_w = ok_l = True
    if width is not None:
        left, right, _, _ = _surrounding(pose, rings)
        if left is not None and right is not None:
            ok_w = not (planar.ring_ring_distance(rings[left], box) + planar.ring_ring_distance(rings[right], box) + c.WIDTH < width)
    if length is not None:
        _, _, front, back = _surrounding(pose, rings)
        if front is not None and back is not None:
            ok_l = not (planar.ring_ring_distance(rings[front], box) + planar.ring_ring_distance(rings[back], box) + c.LENGTH < length)
    return ok_w and ok_l


def _check_extrem_level(start, dest, rings):
    """:121-137"""
    c = _cfg()
    left, right, front, back = _surrounding(dest, rings)
    if math.hypot(start[0] - dest[0], start[1] - dest[1]) > 30.0:
        if front is not None and back is not None and not _has_enough_space(dest, rings, length=c.MIN_PARK_LOT_LEN_DICT["Normal"]):
            return True
        if left is not None and right is not None and not _has_enough_space(dest, rings, width=c.MIN_PARK_LOT_WIDTH_DICT["Normal"]):
            return True
    extrem_len = min(c.LENGTH * 1.2, c.LENGTH + 0.9)  # :11
    if front is not None and back is not None and not _has_enough_space(dest, rings, length=extrem_len):
        return True
    return False


def _free_space_valid(key_pts, rings, skip):
    rect = planar.minimum_rotated_rectangle(key_pts)
    valid = True
    for k, ring in enumerate(rings):
        if any(planar.rings_equal(ring, rings[s]) for s in skip):
            continue
        if planar.convex_polygon_intersects_ring(rect, ring):
            valid = False
    return valid


def get_map_level(start, dest, obstacle_list):
    """map_level.py:27-112 -> "Normal" | "Complex" | "Extrem" """
    c = _cfg()
    if len(obstacle_list) <= 1:
        return LEVEL_NORMAL
    rings = [_ring_of(o) for o in obstacle_list]
    start, dest = _pose(start), _pose(dest)
    if _check_extrem_level(start, dest, rings):
        return LEVEL_EXTREM
    distance_exceed = math.hypot(start[0] - dest[0], start[1] - dest[1]) > c.MAX_DRIVE_DISTANCE
    left, right, front, back = _surrounding(dest, rings)
    rb, rf, lf, lb = _box(dest)
    if left is not None and right is not None and front is None:  # bay parking
        if distance_exceed or not _has_enough_space(dest, rings, width=c.MIN_PARK_LOT_WIDTH_DICT["Normal"]):
            return LEVEL_COMPLEX
        h = dest[2]
        reach = c.BAY_PARK_WALL_DIST_DICT["Normal"] - 0.5
        key = [_translate(lf, h, 0.2), _translate(rf, h, 0.2), _translate(lf, h, reach), _translate(rf, h, reach), (start[0], start[1])]
        return LEVEL_NORMAL if _free_space_valid(key, rings, (left, right)) else LEVEL_COMPLEX
    if front is not None and back is not None:  # parallel parking
        if distance_exceed or not _has_enough_space(dest, rings, length=c.MIN_PARK_LOT_LEN_DICT["Normal"]):
            return LEVEL_COMPLEX
        out = dest[2] + math.pi / 2
        if math.cos(out) * (start[0] - dest[0]) + math.sin(out) * (start[1] - dest[1]) < 0:
            out += math.pi
            k_front, k_back = rf, rb
        else:
            k_front, k_back = lf, lb
        reach = c.PARA_PARK_WALL_DIST_DICT["Normal"] - 0.5
        key = [_translate(k_front, out, 0.2), _translate(k_back, out, 0.2), _translate(k_front, out, reach), _translate(k_back, out, reach)]
        key += _box(start) + [(start[0], start[1])]
        return LEVEL_NORMAL if _free_space_valid(key, rings, (back, front)) else LEVEL_COMPLEX
    if (left is None or right is None) and (front is None or back is None):
        return LEVEL_NORMAL
    return LEVEL_COMPLEX
