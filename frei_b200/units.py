"""
Unit handling at the Python boundary.

The reference speaks ``astropy.units.Quantity`` everywhere.  astropy may be
absent, so every entry point of this package accepts either Quantities (any
convertible unit) or bare numbers/arrays, which are taken to be in the units
the reference's own grids carry:

    temperature K           (frei/tp.py:61)
    pressure    bar         (frei/tp.py:32)
    wavelength  micron      (frei/core.py:39)
    gravity     m s^-2      (the unit named by frei/core.py:68-70)
    mass        g           (frei/core.py:68-70)
    flux        erg s^-1 cm^-3   (frei/twostream.py:13)
    opacity     cm^2 g^-1   (frei/opacity.py:269)

Results are returned as Quantities when astropy is importable and as bare
arrays in the units above otherwise.
"""
import numpy as np

try:                                    # pragma: no cover - astropy is optional
    import astropy.units as _u
    HAVE_ASTROPY = True
except ImportError:
    _u = None
    HAVE_ASTROPY = False

_OPTIONAL = {}


def optional_module(name):
    """Import an optional dependency once per process (a failed import walks sys.path: ~0.1 ms, which is
    1 % of a whole C2 solve when it happens three times per call); None when it is not installed."""
    if name not in _OPTIONAL:
        import importlib
        try:
            _OPTIONAL[name] = importlib.import_module(name)
        except ImportError:
            _OPTIONAL[name] = None
    return _OPTIONAL[name]


# CODATA 2018 (astropy >= 4.0), CGS
m_p = 1.67262192369e-24     # g
amu = 1.66053906660e-24     # g
k_B = 1.380649e-16          # erg / K
sigma_sb = 5.6703744191844314e-5
GM_jup = 1.2668653e23       # cm^3 s^-2
R_jup = 7.1492e9            # cm
au = 1.495978707e13         # cm
R_sun = 6.957e10            # cm

_UNIT_STR = {
    'K': 'K', 'bar': 'bar', 'um': 'um', 'cm': 'cm', 'g': 'g', 'm/s2': 'm / s2',
    'cm/s2': 'cm / s2', 'flux': 'erg / (s cm3)', 'kappa': 'cm2 / g', '': '',
}


def _unit(name):
    return _u.Unit(_UNIT_STR[name])


def is_quantity(x):
    return hasattr(x, 'unit') and hasattr(x, 'to')


def value(x, name):
    """Bare float64 value(s) of ``x`` in unit ``name`` (bare input is assumed to be in it)."""
    if is_quantity(x):
        if name == '':
            return np.asarray(x.to(_u.dimensionless_unscaled).value, dtype=np.float64)
        return np.asarray(x.to(_unit(name)).value, dtype=np.float64)
    if hasattr(x, 'values') and not isinstance(x, dict):      # xarray / pandas
        x = x.values
    return np.asarray(x, dtype=np.float64)


def gravity_cgs(g):
    """Surface gravity in cm s^-2 (bare numbers are m s^-2, see module docstring)."""
    if is_quantity(g):
        return float(g.to(_unit('cm/s2')).value)
    return float(g) * 100.0


def wrap(x, name):
    """Attach unit ``name`` when astropy is available."""
    if HAVE_ASTROPY and name:
        return x * _unit(name)
    return x
