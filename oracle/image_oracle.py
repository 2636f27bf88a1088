"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's image observation:

  CarParking._render                 car_parking_base.py:301-319   painter's order: background, obstacles, start box
                                                                  outline, dest box, vehicle box, last <= 20 trajectory boxes
  CarParking._get_img_observation    car_parking_base.py:321-350   rotate the 500x500 screen by the heading, re-centre on
                                                                  the vehicle box centroid, crop 256x256
  Obs_Processor.process_img          observation_processor.py:13-23  white -> black, cv2.resize to 64x64 (INTER_LINEAR), /255
  observation_rescale                env_wrapper.py:52-55          HWC -> CHW
  Vehicle.reset/step/retreat + the trajectory pruning of CarParking.step   vehicle.py:121-157, car_parking_base.py:273-275

The raster operations are those of oracle/softraster.py (restated pygame; PARITY UNPINNED against a real pygame
build, see its header).  The glue above IS pinned: tests/golden/images_*.npz were recorded from the unmodified
reference running on the same raster restatement and the real cv2 (oracle/make_golden.py --only images), and
tests/test_image_oracle.py replays them through this file.  The 4x INTER_LINEAR down-sampling is restated as
integer arithmetic and checked against the installed cv2 in the same test file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import geom as _g  # noqa: E402
import softraster as sr  # noqa: E402

# src/configs.py:13-24, 80-94, 103
WHEEL_BASE, FRONT_HANG, REAR_HANG, WIDTH = 2.8, 0.96, 0.93, 1.94
VEHICLE_BOX = [(-REAR_HANG, -WIDTH / 2), (FRONT_HANG + WHEEL_BASE, -WIDTH / 2), (FRONT_HANG + WHEEL_BASE, WIDTH / 2),
               (-REAR_HANG, WIDTH / 2)]
BG_COLOR = (255, 255, 255)
START_COLOR = (100, 149, 237)
DEST_COLOR = (69, 139, 0)
OBSTACLE_COLOR = (150, 150, 150)
VEHICLE_COLOR = (30, 144, 255)  # COLOR_POOL[0], vehicle.py:117
TRAJ_RENDER_LEN = 20
TRAJ_COLORS = [tuple(int(v) for v in c[:3]) for c in
               np.linspace(np.array((10, 10, 10, 255)), np.array((10, 10, 200, 255)), TRAJ_RENDER_LEN, endpoint=True, dtype=np.uint8)]
WIN_W = WIN_H = 500
OBS_W = OBS_H = 256
K = 12
DOWNSAMPLE = 4


def create_box(x, y, heading):
    """State.create_box (vehicle.py:32-36): closed ring of 5 coordinates, a*x + b*y + xoff left to right."""
    c, s = np.cos(np.float64(heading)), np.sin(np.float64(heading))
    pts = [(float(c * px + (-s) * py + x), float(s * px + c * py + y)) for px, py in VEHICLE_BOX]
    return pts + [pts[0]]


def screen_matrix(bounds):
    """coord_transform_matrix (car_parking_base.py:140-147)."""
    xmin, xmax, ymin, ymax = (float(b) for b in bounds)
    return K, 0.5 * (WIN_W - K * (xmax + xmin)), 0.5 * (WIN_H - K * (ymax + ymin))


def to_screen(ring, mat):
    """_coord_transform (:149-151): affine [k, 0, 0, k, bx, by] evaluated as a*x + b*y + xoff."""
    k, bx, by = mat
    return [(k * x + 0 * y + bx, 0 * x + k * y + by) for x, y in ring]


def downsample4(img):
    """cv2.resize(img, (w/4, h/4)) with the default INTER_LINEAR, for uint8 HxWx3: the sample point of output
    pixel i is source coordinate 4i + 1.5, i.e. the mean of pixels 4i+1 and 4i+2 with weights 1024/2048 in cv2's
    fixed-point path; both passes together reduce to (a + b + c + d + 2) >> 2."""
    v = img.astype(np.int32)
    s = v[1::4, 1::4] + v[1::4, 2::4] + v[2::4, 1::4] + v[2::4, 2::4]
    return ((s + 2) >> 2).astype(np.uint8)


def traj_colors(traj_render_len):
    """configs.py:87-88 TRAJ_COLORS for another TRAJ_RENDER_LEN"""
    return [tuple(int(v) for v in c[:3]) for c in
            np.linspace(np.array((10, 10, 10, 255)), np.array((10, 10, 200, 255)), traj_render_len, endpoint=True, dtype=np.uint8)]


def render_screen(start, dest, bounds, obstacles, traj, traj_render_len=TRAJ_RENDER_LEN):
    """_render: the 500 x 500 screen.  `obstacles`: list of open vertex lists; `traj`: list of (x, y, heading),
    the last one being the current state.  traj_render_len: configs.py:86 TRAJ_RENDER_LEN, 0 = RENDER_TRAJ off
    (car_parking_base.py:315)."""
    mat = screen_matrix(bounds)
    surf = sr.Surface((WIN_W, WIN_H))
    surf.fill(BG_COLOR)
    for ring in obstacles:
        ring = [tuple(map(float, p)) for p in ring]
        sr.polygon(surf, OBSTACLE_COLOR, to_screen(ring + [ring[0]], mat))
    sr.polygon(surf, START_COLOR, to_screen(create_box(*start), mat), width=1)
    sr.polygon(surf, DEST_COLOR, to_screen(create_box(*dest), mat))
    sr.polygon(surf, VEHICLE_COLOR, to_screen(create_box(*traj[-1]), mat))
    if traj_render_len > 0 and len(traj) > 1:
        n = min(len(traj), traj_render_len)
        colors = TRAJ_COLORS if traj_render_len == TRAJ_RENDER_LEN else traj_colors(traj_render_len)
        for i in range(n):
            sr.polygon(surf, colors[-(n - i)], to_screen(create_box(*traj[-(n - i)]), mat))
    return surf, mat


def crop_observation(surf, mat, pose):
    """_get_img_observation: (256, 256, 3) uint8."""
    x, y, heading = pose
    angle = np.float64(heading)
    old_center = (WIN_W // 2, WIN_H // 2)
    capture = sr.rotate(surf, np.rad2deg(angle))
    rot = sr.Surface((WIN_W, WIN_H))
    rot.blit(capture, capture.get_rect(center=old_center))
    (cx, cy), = to_screen([_g.ring_centroid(create_box(x, y, heading))], mat)
    dx = (cx - old_center[0]) * np.cos(angle) + (cy - old_center[1]) * np.sin(angle)
    dy = -(cx - old_center[0]) * np.sin(angle) + (cy - old_center[1]) * np.cos(angle)
    obs = sr.Surface((WIN_W, WIN_H))
    obs.fill(BG_COLOR)
    obs.blit(rot, (int(-dx), int(-dy)))
    x0, y0 = int((WIN_W - OBS_W) / 2), int((WIN_H - OBS_H) / 2)
    return obs.arr[y0:y0 + OBS_H, x0:x0 + OBS_W].copy()


def process(raw):
    """Obs_Processor.process_img without the final /255.0 (kept as uint8), then HWC -> CHW."""
    img = raw.copy()
    bg = (img == np.array(BG_COLOR, dtype=np.uint8)).sum(axis=-1) == 3
    img[bg] = 0
    return np.ascontiguousarray(downsample4(img).transpose(2, 0, 1))


def render_observation(start, dest, bounds, obstacles, traj, traj_render_len=TRAJ_RENDER_LEN):
    """uint8 (3, 64, 64); the reference's float64 observation is this / 255.0."""
    surf, mat = render_screen(start, dest, bounds, obstacles, traj, traj_render_len)
    return process(crop_observation(surf, mat, traj[-1]))


class TrajectoryBook(object):
    """Vehicle.trajectory as CarParking.step leaves it (one list per env): reset -> [start]; a step appends
    its final state unless the very first substep collided (nothing was kept)."""

    def __init__(self, n):
        self.traj = [[] for _ in range(n)]

    def reset(self, i, pose):
        self.traj[i] = [tuple(float(v) for v in pose)]

    def step(self, i, pose, substeps, retreated):
        if int(substeps) - int(retreated) >= 1:
            self.traj[i].append(tuple(float(v) for v in pose))


def scene_rings(obs, nverts):
    """Open vertex lists of a padded scene (obs[MAX_OBS][4][2], nverts[MAX_OBS])."""
    return [[(float(obs[k, j, 0]), float(obs[k, j, 1])) for j in range(int(nverts[k]))] for k in range(len(nverts)) if nverts[k] > 0]
