"""
frei_b200 — B200-native radiative-equilibrium hot path behind frei's Python API.

Same public names as ``frei/__init__.py`` for the path this package covers:
``Planet``, ``Grid``, ``effective_temperature`` (core), ``chemistry``,
``kappa``, ``load_example_opacity`` (opacity), ``pressure_grid``,
``temperature_grid`` (tp), ``propagate_fluxes``, ``emit``, ``absorb``
(twostream).  All arithmetic on the path runs in hand-written CUDA kernels
(sm_100a) reached through the C ABI in ``include/frei_b200.h``.
"""
from .core import *  # noqa
from .chemistry import *  # noqa
from .opacity import *  # noqa
from .tp import *  # noqa
from .interp import *  # noqa
from .twostream import *  # noqa

__version__ = '0.1.0'
