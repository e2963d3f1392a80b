"""
Minimal stand-in for numpy_groupies (absent here, un-vendored in the reference): only what
frei/interp.py imports to build its own numba Trapz aggregation (frei/interp.py:4-13, 216-243).
The aggregation loop itself (AggregateTrapz._loop, frei/interp.py:174-194) is the reference's
code and runs under the real numba.
"""
from . import utils, utils_numpy, aggregate_numba  # noqa: F401
