"""
Pressure and temperature grids — same interface as ``frei/tp.py``.
"""
import numpy as np

from . import units as U

__all__ = ['pressure_grid', 'temperature_grid']


def pressure_grid(n_layers=30, P_toa=-6, P_boa=1.1):
    """
    Log-spaced pressures from the bottom to the top of the atmosphere
    (frei/tp.py:10-33).  ``P_toa`` / ``P_boa`` are log10(bar).  Returns bar.
    """
    return U.wrap(np.logspace(P_toa, P_boa, n_layers)[::-1], 'bar')


def temperature_grid(pressures, T_ref=2300, P_ref=0.1, alpha=0.1):
    """
    Power-law initial temperature at each pressure: T_ref (P / P_ref)^alpha
    (frei/tp.py:36-62).  Returns K.
    """
    P = U.value(pressures, 'bar')
    return U.wrap(float(U.value(T_ref, 'K')) * (P / float(U.value(P_ref, 'bar'))) ** alpha, 'K')
