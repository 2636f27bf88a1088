"""pygame stand-in (TEST INFRASTRUCTURE ONLY): the software surface of oracle/softraster.py behind the
module layout the reference touches (`car_parking_base.py:301-350, 383-411`), so the UNMODIFIED
reference renders its image observation here.  Display and clock are no-ops."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import softraster as _sr  # oracle/softraster.py

SHOWN = 0
HIDDEN = 1
Surface = _sr.Surface
Rect = _sr.Rect


class _Display(object):
    def init(self):
        pass

    def set_mode(self, size, flags=0):
        return Surface(size)

    def update(self):
        pass

    def quit(self):
        pass


class _Draw(object):
    def polygon(self, surface, color, points, width=0):
        _sr.polygon(surface, color, points, width)


class _Transform(object):
    def rotate(self, surface, angle):
        return _sr.rotate(surface, angle)


class _Image(object):
    def tostring(self, surface, fmt):
        return _sr.tostring(surface, fmt)


class _Clock(object):
    def tick(self, fps=0):
        return 0


class _Time(object):
    Clock = _Clock


display = _Display()
draw = _Draw()
transform = _Transform()
image = _Image()
time = _Time()


def init():
    pass


def quit():
    pass
