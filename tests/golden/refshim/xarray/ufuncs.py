"""xarray.ufuncs stand-in: the reference only calls these on `.values` ndarrays
(state/ensemble.py:160-163, :262-266), where xarray.ufuncs dispatch to numpy."""
from numpy import hypot, sin, cos, radians, arctan2, sqrt  # noqa: F401
