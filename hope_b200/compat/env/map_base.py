"""env/map_base.py:4-16: the `Area(shape, subtype, color)` container of a scene's obstacles."""
import numpy as np


class Area(object):
    def __init__(self, shape=None, subtype=None, color=None):
        self.shape = shape
        self.subtype = subtype
        self.color = color

    def get_shape(self):
        return np.array(self.shape.coords)
