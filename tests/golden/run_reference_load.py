"""
Golden vectors for the LOAD-TIME path, produced by executing the
reference's OWN files — frei/interp.py (its numba Trapz loop runs under the real numba, its
pandas.cut under the real pandas) and frei/opacity.py (binned_opacity, both branches, and
opacity_dir_to_netcdf) — from /root/reference under the stand-ins of refstubs_load/ (xarray,
numpy_groupies: neither installable here) and refstubs/ (astropy).  Writes
tests/golden/reference_load.json:

  F  groupby_bins_agg(array[T, P, n], wavelength, wl_bins, func=np.trapz)          (interp.py:270-307)
  G  binned_opacity(..., groupies=True) and (..., groupies=False) on a synthetic
     line-by-line "netCDF" file                                                     (opacity.py:66-167)
  H  opacity_dir_to_netcdf on synthetic HELIOS-K .bin files, up to the netCDF write  (opacity.py:395-483)
     incl. the single-pressure species that is mirrored to 1/P

    python tests/golden/run_reference_load.py       # only works where /root/reference exists

The stubs are ours: what this pins is the reference's own arithmetic, call order and conventions
(bin membership, trapezoid form, factors, nearest/linear regridding order, file-name parsing,
sample order), with xarray's interpolation semantics restated in refstubs_load/xarray.py.
"""
import importlib
import json
import os
import sys
import tempfile
import types

os.environ.setdefault('TQDM_DISABLE', '1')
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FREI_REFERENCE', '/root/reference')


def load_reference():
    sys.path.insert(0, os.path.join(HERE, 'refstubs'))
    sys.path.insert(0, os.path.join(HERE, 'refstubs_load'))
    pkg = types.ModuleType('frei')
    pkg.__path__ = [os.path.join(REF, 'frei')]
    sys.modules['frei'] = pkg
    mod = types.ModuleType('frei.phoenix')
    mod.get_binned_phoenix_spectrum = None
    sys.modules['frei.phoenix'] = mod
    return {m: importlib.import_module(f'frei.{m}') for m in ('tp', 'chemistry', 'interp', 'opacity', 'core')}


def synthetic_lines(rs, n, nT=3, nP=2):
    """A small line-by-line opacity cube on an irregular ascending wavelength grid [micron]."""
    wl = np.sort(rs.uniform(0.4, 11.0, n))
    wl[n // 3] = wl[n // 3 - 1]                       # a duplicated wavelength sample
    T = np.array([800.0, 1600.0, 2600.0])[:nT]
    P = np.array([1e-3, 10.0])[:nP]
    op = 10 ** rs.uniform(-4, 2, (nT, nP, n)) * (1 + 0.3 * np.sin(40 * wl))
    return wl, T, P, op


def main():
    import warnings
    warnings.simplefilter('ignore')
    m = load_reference()
    import astropy.units as u
    import xarray as xr
    core, opacity, interp = m['core'], m['opacity'], m['interp']
    out = {}
    rs = np.random.RandomState(11)

    planet = core.Planet.from_hot_jupiter()
    grid = core.Grid(planet=planet, T_ref=2400 * u.K, n_layers=6, n_wl_bins=40)
    wl_bins = grid.wl_bins
    wlb = np.asarray(getattr(wl_bins, 'value', wl_bins), dtype=float)

    # ---- case F: groupby_bins_agg with func=np.trapz, as called at frei/opacity.py:137-139 ----
    wl, T, P, op = synthetic_lines(rs, 3000)
    keep = (wl > wlb.min()) & (wl < wlb.max())
    da = xr.DataArray(op[..., keep], dims=['temperature', 'pressure', 'wavelength'],
                      coords=dict(temperature=T, pressure=P, wavelength=wl[keep]), name='opacity')
    res = interp.groupby_bins_agg(da, da.wavelength, wl_bins, func=np.trapz)
    out['F'] = dict(wl=wl.tolist(), T=T.tolist(), P=P.tolist(), opacity=op.tolist(),
                    wl_bins=wlb.tolist(),
                    binned=np.asarray(res.values).tolist(), dims=list(res.dims),
                    centres=np.asarray(res.coords['wavelength']).tolist())
    # float32 samples keep their dtype through the aggregation (check_dtype)
    res32 = interp.groupby_bins_agg(xr.DataArray(op[..., keep].astype(np.float32), dims=da.dims,
                                                 coords=dict(da.coords), name='opacity'),
                                    da.wavelength, wl_bins, func=np.trapz)
    out['F']['binned_f32'] = np.asarray(res32.values, dtype=np.float64).tolist()
    out['F']['binned_f32_dtype'] = str(res32.values.dtype)

    # ---- case G: binned_opacity, both branches, from a synthetic line-by-line file ----
    with tempfile.TemporaryDirectory() as tmp:
        wl2, T2, P2, op2 = synthetic_lines(rs, 2500)
        ds = xr.Dataset(data_vars=dict(opacity=(['temperature', 'pressure', 'wavelength'], op2)),
                        coords=dict(temperature=(['temperature'], T2), pressure=(['pressure'], P2),
                                    wavelength=wl2))
        ds.save_npz(os.path.join(tmp, '1H2-16O__synthetic.nc'))
        g = {}
        for groupies in (True, False):
            r = opacity.binned_opacity(grid.init_temperatures, grid.pressures, grid.wl_bins, grid.lam,
                                       groupies=groupies, species=['H2O'], path=os.path.join(tmp, '*.nc'))
            t = r['1H2-16O']
            g['groupies' if groupies else 'exact'] = dict(
                dims=list(t.dims), values=np.asarray(t.values).tolist(),
                wavelength=np.asarray(t.coords['wavelength']).tolist(),
                temperature=np.asarray(t.coords['temperature']).tolist(),
                pressure=np.asarray(t.coords['pressure']).tolist())
        out['G'] = dict(wl=wl2.tolist(), T=T2.tolist(), P=P2.tolist(), opacity=op2.tolist(),
                        wl_bins=wlb.tolist(),
                        lam=np.asarray(grid.lam.value).tolist(),
                        grid_T=np.asarray(grid.init_temperatures.value).tolist(),
                        grid_P=np.asarray(grid.pressures.to(u.bar).value).tolist(), **g)

    # ---- case H: opacity_dir_to_netcdf on synthetic HELIOS-K .bin files ----
    out['H'] = {}
    for tag, pressures in (('grid', ['n300', 'p000', 'p100']), ('single_pressure', ['p100'])):
        with tempfile.TemporaryDirectory() as tmp:
            d = os.path.join(tmp, '1H2-16O__synthetic')
            os.makedirs(d)
            files = {}
            for T_ in (500, 1500):
                for ptag in pressures:
                    n = len(np.arange(100, 103, 0.01))
                    data = (10 ** rs.uniform(-6, 1, n)).astype(np.float32)
                    name = f'Out_00100_00103_{T_:05d}_{ptag}.bin'
                    data.tofile(os.path.join(d, name))
                    files[name] = data.astype(np.float64).tolist()
            xr.Dataset.written.clear()
            opacity.opacity_dir_to_netcdf(d, os.path.join(tmp, 'out', 'x.nc'))
            (path, ds, enc), = xr.Dataset.written
            op_ = ds.data_vars['opacity']
            out['H'][tag] = dict(files=files, dims=list(op_.dims),
                                 temperature=np.asarray(ds.coords['temperature']).tolist(),
                                 pressure=np.asarray(ds.coords['pressure']).tolist(),
                                 wavelength=np.asarray(ds.coords['wavelength']).tolist(),
                                 opacity=np.asarray(op_.values, dtype=np.float64).tolist(),
                                 opacity_dtype=str(op_.values.dtype))

    with open(os.path.join(HERE, 'reference_load.json'), 'w') as fh:
        json.dump(out, fh)
    print('F binned', np.asarray(out['F']['binned']).shape, '| G', {k: np.asarray(out['G'][k]['values']).shape
                                                                    for k in ('groupies', 'exact')},
          '| H', {k: np.asarray(v['opacity']).shape for k, v in out['H'].items()})


if __name__ == '__main__':
    main()
