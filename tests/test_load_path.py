"""
Load-time path against outputs of the reference's OWN code:
tests/golden/reference_load.json was produced by tests/golden/run_reference_load.py, which runs
frei/interp.py (numba Trapz loop, pandas.cut) and frei/opacity.py (binned_opacity, both branches;
opacity_dir_to_netcdf) from the reference checkout under dependency stand-ins.

CPU tests pin the oracle (oracle/binning_oracle.py) and the host-side file reader to those
vectors; GPU tests compare the CUDA path (csrc/binning.cu through the C ABI) with them.
"""
import json
import os

import numpy as np
import pytest

from oracle import binning_oracle as BO

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def gold():
    with open(os.path.join(HERE, 'golden', 'reference_load.json')) as fh:
        return json.load(fh)


def _case_F(gold):
    F = gold['F']
    wl, bins, op = np.array(F['wl']), np.array(F['wl_bins']), np.array(F['opacity'])
    keep = (wl > bins.min()) & (wl < bins.max())
    return F, wl[keep], bins, op[..., keep]


def _case_G(gold):
    G = gold['G']
    return G, {k: np.array(G[k]) for k in ('wl', 'T', 'P', 'opacity', 'wl_bins', 'lam', 'grid_T', 'grid_P')}


# ---------------------------------------------------------------------------------------------
# CPU: oracle and host code vs the reference's own outputs
# ---------------------------------------------------------------------------------------------
def test_oracle_groupby_bins_agg_equals_reference_run(gold):
    F, wl, bins, op = _case_F(gold)
    out, centres = BO.groupby_bins_agg(op, wl, bins)
    assert np.array_equal(out, np.array(F['binned']))                       # same additions, same order
    assert np.array_equal(centres, np.array(F['centres']))
    assert F['dims'] == ['temperature', 'pressure', 'wavelength']
    # float32 samples: the reference accumulates in float32 (numpy_groupies keeps the input type)
    assert F['binned_f32_dtype'] == 'float32'
    out32, _ = BO.groupby_bins_agg(op.astype(np.float32), wl, bins, keep_dtype=True)
    assert np.array_equal(out32.astype(np.float64), np.array(F['binned_f32']))


def test_oracle_binned_opacity_both_branches_equal_reference_run(gold):
    G, a = _case_G(gold)
    tab, centres = BO.binned_opacity_one(a['opacity'], a['wl'], a['T'], a['P'], a['grid_T'], a['grid_P'], a['wl_bins'])
    g = G['groupies']
    assert g['dims'] == ['temperature', 'pressure', 'wavelength']
    assert np.array_equal(tab, np.array(g['values']))
    assert np.array_equal(centres, np.array(g['wavelength']))
    assert np.array_equal(np.array(g['temperature']), a['grid_T']) and np.array_equal(np.array(g['pressure']), a['grid_P'])
    ex = BO.binned_opacity_exact_one(a['opacity'], a['wl'], a['T'], a['P'], a['grid_T'], a['grid_P'],
                                     a['wl_bins'], a['lam'])
    e = G['exact']
    assert e['dims'] == ['wavelength', 'temperature', 'pressure']
    np.testing.assert_allclose(ex, np.array(e['values']), rtol=1e-14)
    assert np.array_equal(np.array(e['wavelength']), a['lam'])
    # the two branches are different algorithms with different normalisation
    assert not np.allclose(np.transpose(ex, (1, 2, 0)), tab, rtol=0.2)


@pytest.mark.parametrize('tag', ['grid', 'single_pressure'])
def test_read_opacity_dir_equals_reference_opacity_dir_to_netcdf(gold, tag, tmp_path):
    """frei/opacity.py:395-483 up to the netCDF write, incl. the single-pressure mirror to 1/P."""
    from frei_b200.opacity import read_opacity_dir
    H = gold['H'][tag]
    d = tmp_path / '1H2-16O__synthetic'
    d.mkdir()
    for name, data in H['files'].items():
        np.asarray(data, dtype=np.float32).tofile(str(d / name))
    T, P, wl, grid = read_opacity_dir(str(d))
    assert H['dims'] == ['temperature', 'pressure', 'wavelength'] and H['opacity_dtype'] == 'float32'
    assert np.array_equal(T, np.array(H['temperature'], dtype=float))
    assert np.array_equal(P, np.array(H['pressure']))
    assert np.array_equal(wl, np.array(H['wavelength']))
    assert grid.dtype == np.float32 and np.array_equal(grid.astype(np.float64), np.array(H['opacity']))


def test_nearest_index_is_scipy_interp1d_nearest():
    from scipy.interpolate import interp1d
    from frei_b200.opacity import nearest_index
    rs = np.random.RandomState(3)
    for axis in (np.array([800.0, 1600.0, 2600.0]), np.array([10.0, 0.1]), rs.uniform(0, 10, 17)):
        q = np.concatenate([rs.uniform(axis.min() - 5, axis.max() + 5, 200), axis,
                            0.5 * (np.sort(axis)[1:] + np.sort(axis)[:-1])])           # incl. exact ties
        ref = interp1d(axis, np.arange(len(axis)), kind='nearest', fill_value='extrapolate',
                       assume_sorted=False)(q).astype(int)
        assert np.array_equal(nearest_index(axis, q), ref)


# ---------------------------------------------------------------------------------------------
# GPU: the CUDA path vs the reference's own outputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_groupby_bins_agg_vs_reference_run(gold):
    from frei_b200.interp import groupby_bins_agg
    F, wl, bins, op = _case_F(gold)
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    out = groupby_bins_agg(op, wl, bins, func=trapz)
    np.testing.assert_allclose(out, np.array(F['binned']), rtol=1e-12)
    assert np.array_equal(out.wavelength, np.array(F['centres']))
    # float32 samples: the kernel accumulates in fp64, the reference in float32 (75 samples per
    # bin here: its own rounding noise is ~1e-6); against the fp64 sum of the same samples 1e-12
    out32 = groupby_bins_agg(op.astype(np.float32), wl, bins)
    np.testing.assert_allclose(out32, np.array(F['binned_f32']), rtol=2e-5)
    exact32, _ = BO.groupby_bins_agg(op.astype(np.float32), wl, bins)
    np.testing.assert_allclose(out32, exact32, rtol=1e-12)


@pytest.mark.gpu
def test_gpu_bin_and_regrid_both_branches_vs_reference_run(gold):
    from frei_b200.opacity import bin_and_regrid
    G, a = _case_G(gold)
    for groupies, key in ((True, 'groupies'), (False, 'exact')):
        tab = bin_and_regrid(a['opacity'], a['wl'], a['T'], a['P'], a['grid_T'], a['grid_P'], a['wl_bins'],
                             lam=a['lam'], groupies=groupies)
        assert tab.dims == ('pressure', 'temperature', 'wavelength')
        ref = np.array(G[key]['values'])
        ref = np.transpose(ref, (1, 0, 2)) if groupies else np.transpose(ref, (2, 1, 0))   # -> [P, T, wl]
        np.testing.assert_allclose(tab.values, ref, rtol=1e-12)
        assert np.array_equal(tab.wavelength, np.array(G[key]['wavelength']))
        assert np.array_equal(tab.temperature, a['grid_T']) and np.array_equal(tab.pressure, a['grid_P'])


@pytest.mark.gpu
def test_gpu_binned_opacity_from_bin_directories(tmp_path):
    """binned_opacity end to end (.bin directory -> binned tables) against the oracle chain, both
    branches, float32 files as HELIOS-K writes them; a bin with a single sample gives NaN in the
    groupies=False branch exactly where the oracle (= the reference's 0/0) has it."""
    from frei_b200.opacity import binned_opacity, read_opacity_dir
    import frei_b200 as frei
    rs = np.random.RandomState(5)
    d = tmp_path / '1H2-16O__x'
    d.mkdir()
    n = len(np.arange(1000, 1400, 0.01))                   # 7.14 .. 10 micron
    for T_ in (600, 1800, 3000):
        for ptag in ('n200', 'p000', 'p200'):
            (10 ** rs.uniform(-5, 1, n)).astype(np.float32).tofile(str(d / f'Out_01000_01400_{T_:05d}_{ptag}.bin'))
    grid = frei.Grid(frei.Planet.from_hot_jupiter(), n_wl_bins=60, n_layers=7, lam_min=7.5, lam_max=9.5)
    T, P, wl, cube = read_opacity_dir(str(d))
    from frei_b200 import units as U
    gT, gP = U.value(grid.init_temperatures, 'K'), U.value(grid.pressures, 'bar')
    bins, lam = U.value(grid.wl_bins, 'um'), U.value(grid.lam, 'um')
    for groupies in (True, False):
        tabs = binned_opacity(grid.init_temperatures, grid.pressures, grid.wl_bins, grid.lam,
                              groupies=groupies, species=['H2O'], path=str(tmp_path / '*'))
        got = tabs['1H2-16O'].values                       # [P, T, wl]
        if groupies:
            ref, _ = BO.binned_opacity_one(cube, wl, T, P, gT, gP, bins)
            ref = np.transpose(ref, (1, 0, 2))
        else:
            ref = np.transpose(BO.binned_opacity_exact_one(cube, wl, T, P, gT, gP, bins, lam), (2, 1, 0))
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        np.testing.assert_allclose(got, ref, rtol=1e-11, equal_nan=True)
    # single-sample bin -> NaN node -> NaN neighbours after the linear interpolation
    from frei_b200.opacity import bin_and_regrid
    wl_s = np.array([0.50, 0.51, 0.52, 0.75, 0.90, 0.91, 0.92])
    op_s = rs.uniform(1, 2, (2, 2, 7))
    bins_s = np.array([0.4, 0.6, 0.8, 1.0])
    lam_s = np.array([0.5, 0.7, 0.9])
    got = bin_and_regrid(op_s, wl_s, [500.0, 900.0], [0.1, 10.0], [600.0], [1.0], bins_s, lam=lam_s,
                         groupies=False).values
    ref = np.transpose(BO.binned_opacity_exact_one(op_s, wl_s, np.array([500.0, 900.0]), np.array([0.1, 10.0]),
                                                   np.array([600.0]), np.array([1.0]), bins_s, lam_s), (2, 1, 0))
    assert np.isnan(ref).any() and np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=1e-12, equal_nan=True)
