"""Stand-in for the shapely subset the reference env uses. See oracle/refshim/README.md."""
__version__ = "1.8-standin"
