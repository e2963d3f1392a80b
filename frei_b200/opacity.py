"""
Opacity evaluation — same interface as the hot-path part of ``frei/opacity.py``.

``kappa()`` runs on the GPU: brackets/weights (K0) and the gather (K1) of
``csrc/frei_b200.cu``.  There is no CPU implementation in this package.
"""
import numpy as np

from . import units as U
from .chemistry import chemistry

__all__ = ['kappa', 'load_example_opacity', 'OpacityTable', 'binned_opacity']


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


# ---------------------------------------------------------------------------------------------
# Load-time path: HELIOS-K / DACE ".bin" directory -> binned tables
# ---------------------------------------------------------------------------------------------
def _parse_bin_name(filename):
    """(wavenumber start, end, temperature, pressure[bar]) from 'Out_<w0>_<w1>_<T>_<p|n><100 log10 P>.bin'
    (frei/opacity.py:403-410)."""
    f = filename.split('_')
    sign = 1 if f[4][0] == 'p' else -1
    return int(f[1]), int(f[2]), int(f[3]), 10 ** (sign * float(f[4][1:].split('.')[0]) / 100)


def read_opacity_dir(opacity_dir):
    """
    Read a directory of HELIOS-K float32 ``.bin`` cross-section files into
    ``(temperature[T], pressure[P], wavelength_um[N] ascending, opacity[T, P, N] float32)`` —
    what ``opacity_dir_to_netcdf`` (frei/opacity.py:395-483) stores in its netCDF file: first
    sample dropped, order reversed to ascending wavelength, 0.01 cm^-1 wavenumber grid, and a
    species with a single pressure duplicated at 1/P.
    """
    import os
    files = [(dp, fn) for dp, _, fns in os.walk(opacity_dir) for fn in sorted(fns) if fn.endswith('.bin')]
    if not files:
        raise FileNotFoundError(f'no .bin files under {opacity_dir}')
    meta = [_parse_bin_name(fn) for _, fn in files]
    w0, w1 = meta[-1][0], meta[-1][1]
    wavelength = (1 / np.arange(w0, w1, 0.01) / 1e-4)[1:][::-1]
    tgrid = np.sort(list({m[2] for m in meta}))
    pgrid = np.sort(list({m[3] for m in meta}))
    mirror = len(pgrid) == 1
    if mirror:
        pgrid = np.concatenate([pgrid, 10 ** (-1 * np.log10(pgrid))])
    grid = np.zeros((len(tgrid), len(pgrid), len(wavelength)), dtype='float32')
    for flip in ([False, True] if mirror else [False]):
        for (dp, fn), (_, _, T, P) in zip(files, meta):
            data = np.fromfile(os.path.join(dp, fn), dtype=np.float32)[1:][::-1]
            if flip:
                P = 1.0 / P
            grid[np.argmin(np.abs(tgrid - T)), np.argmin(np.abs(pgrid - P)), :] = data
    return tgrid.astype(np.float64), pgrid, wavelength, grid


def nearest_index(axis, query):
    """Index of the nearest node, ties to the lower node, clipped at the ends — scipy's
    ``interp1d(kind='nearest', fill_value='extrapolate')`` as xarray uses it in
    frei/opacity.py:141-146."""
    axis = np.asarray(axis, dtype=np.float64)
    order = np.argsort(axis, kind='stable')
    mids = 0.5 * (axis[order][1:] + axis[order][:-1])
    return order[np.searchsorted(mids, np.asarray(query, dtype=np.float64), side='left')]


def _regrid_device(binned, src_T, src_P, temperatures, pressures, lerp=None):
    """
    ``frei_b200_regrid``: nearest-neighbour (T, P) lookup of the Grid's nodes in the source grid
    and (``lerp = (x_src, x_new)``) linear interpolation with extrapolation along wavelength, both
    in scipy's interp1d arithmetic as xarray applies it (frei/opacity.py:141-146, 163-166).
    binned: CUDA tensor [nT][nP][nb] fp64.  Returns a CUDA tensor [mT][mP][m].
    """
    import torch
    from . import _cabi
    lib = _cabi.load()
    dev = binned.device
    nT, nP, nb = binned.shape
    iT = torch.from_numpy(nearest_index(src_T, temperatures).astype(np.int32)).to(dev)
    iP = torch.from_numpy(nearest_index(src_P, pressures).astype(np.int32)).to(dev)
    j0 = dx = t = None
    m = nb
    if lerp is not None:
        x, x_new = (np.asarray(v, dtype=np.float64) for v in lerp)
        if x.shape[0] != nb or nb < 2:
            raise ValueError('linear interpolation along wavelength needs at least two source bins')
        hi = np.clip(np.searchsorted(x, x_new), 1, nb - 1)               # scipy interp1d._call_linear
        lo = hi - 1
        j0 = torch.from_numpy(lo.astype(np.int32)).to(dev)
        dx = torch.from_numpy(x[hi] - x[lo]).to(dev)
        t = torch.from_numpy(x_new - x[lo]).to(dev)
        m = x_new.shape[0]
    out = torch.empty((iT.shape[0], iP.shape[0], m), dtype=torch.float64, device=dev)
    _cabi.check(lib.frei_b200_regrid(binned.contiguous().data_ptr(), nT, nP, nb, iT.data_ptr(), iT.shape[0],
                                     iP.data_ptr(), iP.shape[0], _cabi.ptr(j0), _cabi.ptr(dx), _cabi.ptr(t),
                                     m, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return out


def bin_and_regrid(opacity, wavelength_um, src_T, src_P, temperatures, pressures, wl_bins, lam=None,
                   groupies=True):
    """
    One species of ``binned_opacity`` on the GPU.  ``opacity[T, P, n]`` are line-by-line samples at
    ascending ``wavelength_um``.

    groupies=True (frei/opacity.py:128-146): crop to the bin range, unit-spacing trapezoid sum per
    wavelength bin, times bin width times 1e-3, then nearest-neighbour lookup onto the Grid's
    (T, P); wavelength coordinate = the bin centres of ``pandas.cut``'s labels.

    groupies=False (frei/opacity.py:29-40, 150-167; the default of ``Grid.load_opacities``): per
    bin with at least one sample, the trapezoid integral over wavelength of the nearest (T, P)
    node divided by the wavelength span of the bin's samples, placed at their mean wavelength;
    then linear interpolation with extrapolation onto ``lam``.  A bin with a single sample is 0/0 =
    NaN there, as in the reference.

    Returns an OpacityTable with dims (temperature, pressure, wavelength).
    """
    import torch
    from .interp import bin_trapz_device, bin_centres, cut_codes, _device_samples
    wl_bins = np.asarray(U.value(wl_bins, 'um'), dtype=np.float64)
    wl = np.asarray(wavelength_um, dtype=np.float64)
    T_new, P_new = U.value(temperatures, 'K'), U.value(pressures, 'bar')
    if groupies:
        keep = (wl > wl_bins.min()) & (wl < wl_bins.max())             # :131-135
        a = _device_samples(np.asarray(opacity)[..., keep])
        binned, _ = bin_trapz_device(a, wl[keep], wl_bins)             # [T, P, n_bins]
        width = torch.from_numpy((wl_bins[1:] - wl_bins[:-1]) * 1e-3).to(binned.device)      # :139
        out = _regrid_device(binned * width, src_T, src_P, T_new, P_new)
        centres = bin_centres(wl_bins)
    else:
        if lam is None:
            raise ValueError('groupies=False interpolates onto lam: pass the wavelength grid')
        if np.any(np.diff(wl) < 0):
            raise ValueError('wavelength samples must be ascending')
        codes = cut_codes(wl, wl_bins)
        inside = codes >= 0
        a = _device_samples(np.asarray(opacity)[..., inside])
        integ, first = bin_trapz_device(a, wl[inside], wl_bins, positions=True)
        occupied = np.flatnonzero(np.bincount(codes[inside], minlength=wl_bins.shape[0] - 1))
        # per occupied bin: span and mean of its samples (groupby_bins skips empty bins, :155-160)
        c = codes[inside]
        w_in = wl[inside]
        lo = np.searchsorted(c, occupied, side='left')
        hi = np.searchsorted(c, occupied, side='right')
        span = w_in[hi - 1] - w_in[lo]                                 # wl.max() - wl.min(), :38
        mean = np.array([w_in[s:e].mean() for s, e in zip(lo, hi)])    # wl.mean(), :40
        sel = torch.from_numpy(occupied).to(integ.device)
        binned = integ.index_select(2, sel) / torch.from_numpy(span).to(integ.device)       # 0/0 -> NaN
        out = _regrid_device(binned, src_T, src_P, T_new, P_new, lerp=(mean, U.value(lam, 'um')))
        centres = np.asarray(U.value(lam, 'um'), dtype=np.float64)
    tab = OpacityTable(np.transpose(out.cpu().numpy(), (1, 0, 2)), P_new, T_new, centres)
    tab.dims = ('pressure', 'temperature', 'wavelength')
    return tab


def binned_opacity(temperatures, pressures, wl_bins, lam, groupies=True, species=None, path=None):
    """
    Opacity tables of all available species binned to the Grid's wavelengths — signature of
    frei/opacity.py:66-69, both branches (``groupies``: see :func:`bin_and_regrid`).  ``path`` is a glob of HELIOS-K ``.bin`` directories named
    ``<isotopologue>_*`` (the reference reads the netCDF files it wrote from those directories;
    netCDF is not available here, the ``.bin`` payload is identical).
    """
    import os
    from glob import glob
    from .chemistry import iso_to_species
    if path is None:
        path = os.path.join(os.path.expanduser('~'), '.frei', '*')
    dirs = [p for p in sorted(glob(path)) if os.path.isdir(p)]
    iso = [os.path.basename(p).split('_')[0] for p in dirs]
    if species is not None:
        keep = [iso_to_species(i) in species for i in iso]
        dirs, iso = [d for d, k in zip(dirs, keep) if k], [i for i, k in zip(iso, keep) if k]
    if not dirs:
        raise FileNotFoundError(f'no opacity directories match {path}')
    results = {}
    for name, d in zip(iso, dirs):
        T, P, wl, grid = read_opacity_dir(d)
        results[name] = bin_and_regrid(grid, wl, T, P, temperatures, pressures, wl_bins, lam=lam,
                                       groupies=groupies)
    return results
