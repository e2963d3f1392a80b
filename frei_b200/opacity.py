"""
Opacity evaluation — same interface as the hot-path part of ``frei/opacity.py``.

``kappa()`` runs on the GPU: brackets/weights (K0) and the gather (K1) of
``csrc/frei_b200.cu``.  There is no CPU implementation in this package.
"""
import numpy as np

from . import units as U
from .chemistry import chemistry

__all__ = ['kappa', 'load_example_opacity', 'OpacityTable']


class OpacityTable:
    """
    Minimal stand-in for the ``xarray.DataArray`` opacity tables of the reference
    (dims pressure, temperature, wavelength; coords in bar, K, micron;
    frei/opacity.py:331-339) for when xarray is not installed.
    """
    dims = ('pressure', 'temperature', 'wavelength')

    def __init__(self, values, pressure, temperature, wavelength):
        self.values = np.asarray(values)
        self.pressure = np.asarray(pressure, dtype=np.float64)
        self.temperature = np.asarray(temperature, dtype=np.float64)
        self.wavelength = np.asarray(wavelength, dtype=np.float64)

    @property
    def shape(self):
        return self.values.shape

    def drop_duplicates(self, dim):
        if dim != 'temperature':
            raise NotImplementedError(dim)
        _, first = np.unique(self.temperature, return_index=True)
        keep = np.sort(first)
        return OpacityTable(self.values[:, keep], self.pressure, self.temperature[keep],
                            self.wavelength)


_table_cache = {}


def device_table(opacities, dtype=None, lam_range=None):
    """Upload (once per opacity dict) and return the :class:`~frei_b200.engine.DeviceTable`."""
    from .engine import DeviceTable, FREI_F64
    dtype = FREI_F64 if dtype is None else dtype
    key = (id(opacities), dtype, lam_range)
    hit = _table_cache.get(key)
    if hit is not None and hit[0] is opacities:
        return hit[1]
    tab = DeviceTable(opacities, dtype=dtype, lam_range=lam_range)
    if len(_table_cache) > 8:
        _table_cache.clear()
    _table_cache[key] = (opacities, tab)
    return tab


def kappa(opacities, temperature, pressure, lam, m_bar=2.4 * U.m_p):
    """
    Total opacity k(lambda) = sum_s mmr_s * interp_{P,T}(table_s) + sigma and the
    Rayleigh scattering cross-section sigma(lambda); same signature and return
    order as frei/opacity.py:203-269.  Scalar (T, P) give flat [n_lam] arrays,
    vectors of length n give [n, n_lam].
    """
    from .engine import Engine
    T = np.atleast_1d(U.value(temperature, 'K'))
    P = np.atleast_1d(U.value(pressure, 'bar'))
    lam_um = U.value(lam, 'um')
    m_bar_g = float(U.value(m_bar, 'g'))
    mmr_d = chemistry(T, P, opacities.keys(), m_bar=m_bar_g)
    mmr = np.stack([np.broadcast_to(mmr_d[s], T.shape) for s in opacities], axis=-1)
    n = T.shape[0]
    # the engine wants >= 3 levels; pad by repeating the last point
    pad = max(0, 3 - n)
    Tp = np.concatenate([T, np.repeat(T[-1:], pad)])
    Pp = np.concatenate([P, np.repeat(P[-1:], pad)])
    mp = np.concatenate([mmr, np.repeat(mmr[-1:], pad, axis=0)])
    eng = Engine(device_table(opacities), lam_um, Pp, Tp, mp, g=1.0, m_bar=m_bar_g)
    k, sg = eng.kappa()
    k = k[0, :n].cpu().numpy()
    sg = sg[0].cpu().numpy()
    if n == 1:
        k = k[0]
    return U.wrap(k, 'kappa'), U.wrap(sg, 'kappa')


def load_example_opacity(grid, seed=42, scale_factor=20):
    """
    Synthetic water-like opacity table on the grid's (P, T, lambda), the
    reference's own fixture (frei/opacity.py:272-342): two broad Gaussians, 15
    seeded narrow optical bands (legacy ``np.random.seed`` stream: amplitudes
    drawn before centres) and three near-infrared bands, flat in T and P.
    """
    lam = U.value(grid.lam, 'um')
    P = U.value(grid.pressures, 'bar')
    T = U.value(grid.init_temperatures, 'K')
    np.random.seed(seed)
    amps = np.random.uniform(low=0.1, high=0.2, size=15)
    cens = np.random.uniform(low=0.5, high=1, size=15)
    so = np.exp(-0.5 * (lam - 6) ** 2 / 2 ** 2) + 0.8 * np.exp(-0.5 * (lam - 0.3) ** 2 / 0.5 ** 2)
    for amp, wl in zip(amps, cens):
        so = so + amp * np.exp(-0.5 * (lam - wl) ** 2 / 0.005 ** 2)
    for amp, wl in zip([0.22, 0.2, 0.18], np.logspace(np.log10(1.4), np.log10(2.7), 3)):
        so = so + amp * np.exp(-0.5 * (lam - wl) ** 2 / 0.13 ** 2)
    vals = np.zeros((P.shape[0], T.shape[0], lam.shape[0]))
    vals[:] += 5 * 10 ** (2.5 * (so - 0.4))
    vals *= scale_factor
    try:                                                  # pragma: no cover
        import xarray as xr
        tab = xr.DataArray(vals, dims=['pressure', 'temperature', 'wavelength'],
                           coords=dict(pressure=P, temperature=T, wavelength=lam)
                           ).drop_duplicates('temperature')
    except ImportError:
        tab = OpacityTable(vals, P, T, lam).drop_duplicates('temperature')
    return {"1H2-16O": tab}
