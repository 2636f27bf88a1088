"""shapely.geometry.polygon: the module path data/dlp.data's pickled LinearRings refer to."""
from . import LinearRing, Polygon  # noqa: F401
