"""
CPU tests of the oracle (no GPU): it must reproduce every known-answer value the
reference's own tests hold for the hot path (frei/tests/test_core.py) and the
committed golden vectors (tests/golden), and its explicit bracket/weights must
agree with scipy — the arithmetic xarray dispatches to in frei/opacity.py:261-263.
"""
import json
import os

import numpy as np
import pytest

from oracle import frei_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def default_case():
    """Default Grid of the reference's test: 30 layers, 500 bins, T_ref = 2400 K, scale_factor=1."""
    pl = O.hot_jupiter()
    P = O.pressure_grid(30, np.log10(1e-6), np.log10(200))          # frei/core.py:123
    T = O.temperature_grid(P, 2400.0, 0.1, 0.1)
    lam, _, _ = O.wavelength_grid(0.5, 10, 500)
    tabs = O.load_example_opacity(P, T, lam, scale_factor=1)
    return pl, P, T, lam, tabs


def test_constants_and_grids(default_case):
    pl, P, T, lam, tabs = default_case
    assert abs(pl['g'] - 2478.6519476149147) < 1e-9                  # G M_jup / R_jup^2 in cgs
    assert abs(pl['a_rstar'] - 6.45096467011643) < 1e-12             # 0.03 au / R_sun
    assert P[0] > P[-1] and abs(P[0] - 200) < 1e-9 and abs(P[-1] - 1e-6) < 1e-15
    assert abs(T[0] - 5132.326079957703) < 1e-9 and abs(T[-1] - 758.9466384404109) < 1e-9
    assert tabs['1H2-16O']['values'].shape == (30, 30, 500)


def test_reference_kat_inequalities(default_case):
    """frei/tests/test_core.py:34-44: k > sigma everywhere; sigma decreasing (Rayleigh)."""
    pl, P, T, lam, tabs = default_case
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    k, sigma = O.kappa(tabs, T[0], P[0], lam, mmr, pl['m_bar'])
    assert np.all(k > sigma)
    assert sigma[0] > sigma[-1]


def test_reference_kat_values(default_case):
    """
    frei/tests/test_core.py:46-71: peak wavelength 1.1518 +- 0.02 um, peak flux
    1.296e13 +- 0.1e13 erg/s/cm^3, T_eff 2400 +- 200 K after emission_spectrum(n_timesteps=1).
    The reference's CI runs with pyfastchem, whose H2O abundance at these (T, P) is ~3e-4
    (frei/tests/test_chemistry.py:46); with that mixing ratio the oracle meets all three.
    With the mock's 1.5e-3 (no pyfastchem) the reference's own test does not pass either.
    """
    pl, P, T, lam, tabs = default_case
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'], vmr=3e-4)
    spec, Tf, hist, dtaus, n_it = O.emission_spectrum(tabs, T, P, lam, pl, lambda a, b: mmr,
                                                      n_timesteps=1)
    assert abs(lam[spec.argmax()] - 1.1518) < 0.02
    assert abs(spec.max() - 1.296e13) < 0.1e13
    assert abs(O.effective_temperature(P, lam, spec, dtaus, Tf) - 2400) < 200
    assert dtaus.shape == (30, 500) and np.all(dtaus[0] == 1)         # leading row of ones
    assert hist.shape == (30, 2)


def test_interpolation_matches_scipy_bit_for_bit():
    """Explicit bracket + weights == scipy's find_indices / interpn on and off the nodes."""
    from scipy.interpolate import interpn
    from scipy.interpolate._rgi_cython import find_indices
    rs = np.random.RandomState(0)
    axP = 10.0 ** np.arange(-7, 5)
    axT = 200.0 + 400.0 * np.arange(24)
    vals = rs.uniform(0.1, 10, (12, 24, 40))
    pts_P = np.concatenate([axP, np.nextafter(axP, 0), np.nextafter(axP, 1e9), [1e-9, 1e5],
                            10 ** rs.uniform(-7.5, 4.5, 200)])
    pts_T = np.concatenate([axT, rs.uniform(100, 9600, pts_P.size - axT.size)])
    iP, wP = find_indices((axP,), pts_P[None])
    iT, wT = find_indices((axT,), pts_T[None])
    for n, (p, t) in enumerate(zip(pts_P, pts_T)):
        ip, wp, op = O.bracket(axP, p)
        it, wt, ot = O.bracket(axT, t)
        assert (ip, it) == (iP[0, n], iT[0, n])
        assert wp == wP[0, n] and wt == wT[0, n]
        ref = interpn((axP, axT), vals, np.array([[p, t]]), method='linear', bounds_error=False,
                      fill_value=0)[0]
        tab = {'x': dict(P=axP, T=axT, values=vals)}
        lam = np.linspace(1, 2, 40)
        k, s = O.kappa_explicit(tab, t, p, lam, [1.0])
        k2, s2 = O.kappa(tab, t, p, lam, [1.0])
        assert np.array_equal(k2 - s2, ref) or np.allclose(k2 - s2, ref, rtol=1e-15, atol=0)
        np.testing.assert_allclose(k - s, ref, rtol=1e-14, atol=1e-300)
        if op or ot:
            assert np.all(ref == 0)


def test_pressure_only_tables():
    """A table with a single unique temperature is interpolated in pressure only (opacity.py:256)."""
    axP = np.array([1e-3, 1e-1, 10.0])
    vals = np.arange(3 * 1 * 5, dtype=float).reshape(3, 1, 5) + 1
    tab = {'x': dict(P=axP, T=np.array([1000.0]), values=vals)}
    lam = np.linspace(1, 2, 5)
    k, s = O.kappa(tab, 4321.0, 1.0, lam, [2.0])
    w = (1.0 - 0.1) / (10.0 - 0.1)
    np.testing.assert_allclose(k - s, 2.0 * (vals[1, 0] + w * (vals[2, 0] - vals[1, 0])), rtol=1e-14)
    k, s = O.kappa(tab, 4321.0, 100.0, lam, [2.0])
    assert np.array_equal(k, s)                                      # out of range -> 0


def test_two_stream_E_branches_and_limits():
    w0 = np.array([0.0, 0.05, 0.1, 0.100001, 0.5])
    Ew = O.E(w0)
    assert np.all(Ew[:3] == 1) and abs(Ew[4] - (1.225 - 0.1777 * 0.5 - 0.05582 * 0.25)) < 1e-15
    # pure absorption, isothermal, optically thick: both streams relax to pi B
    lam = np.array([1e-4, 3e-4])
    F2u, F1d = O.propagate_fluxes(lam, np.zeros(2), np.zeros(2), 1500.0, 1500.0,
                                  np.full(2, 50.0), np.zeros(2))
    np.testing.assert_allclose(F2u, np.pi * O.BB(1500.0, lam), rtol=1e-12)
    np.testing.assert_allclose(F1d, np.pi * O.BB(1500.0, lam), rtol=1e-12)
    # transparent layer: fluxes pass through
    F2u, F1d = O.propagate_fluxes(lam, np.array([3e14, 4e14]), np.array([5e14, 6e14]), 1500.0,
                                  1500.0, np.full(2, 1e-12), np.zeros(2))
    np.testing.assert_allclose(F2u, [3e14, 4e14], rtol=1e-9)
    np.testing.assert_allclose(F1d, [5e14, 6e14], rtol=1e-9)


def test_sweep_quirks(default_case):
    """Appendix C of SURVEY.md: F_up[0] stays 0, emit leaves T[0], absorb leaves T[-1], dtaus order."""
    pl, P, T, lam, tabs = default_case
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    F = O.F_TOA(lam * 1e-4, pl['T_star'], a_rstar=pl['a_rstar'])
    Fu, Fd = np.zeros((30, 500)), np.zeros((30, 500))
    e = O.emit(tabs, T, P, lam, F, pl['g'], pl['m_bar'], lambda a, b: mmr, fluxes_up=Fu, fluxes_down=Fd)
    assert e[5][0] == 0 and e[2][0] == T[0] and np.all(Fu[0] == 0)
    a = O.absorb(tabs, e[2], P, lam, F, pl['g'], pl['m_bar'], lambda a, b: mmr, fluxes_up=Fu,
                 fluxes_down=Fd)
    assert a[5][-1] == 0 and a[2][-1] == e[2][-1] and np.all(Fu[0] == 0)
    # absorb's dtaus rows are in visiting order (top -> bottom) after the row of ones
    e2 = O.emit(tabs, a[2], P, lam, F, pl['g'], pl['m_bar'], lambda a_, b: mmr, fluxes_up=Fu.copy(),
                fluxes_down=Fd.copy())
    a2 = O.absorb(tabs, a[2], P, lam, F, pl['g'], pl['m_bar'], lambda a_, b: mmr, fluxes_up=Fu.copy(),
                  fluxes_down=Fd.copy())
    np.testing.assert_allclose(a2[4][1:][::-1][1:], e2[4][1:-1], rtol=1e-13)


def test_convergence_anchor(default_case):
    """Full solve on the default grid stops after 53 iterations (sign-flip / <3 K rule, core.py:306-318)."""
    pl, P, T, lam, tabs = default_case
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'], vmr=3e-4)
    spec, Tf, hist, dtaus, n_it = O.emission_spectrum(tabs, T, P, lam, pl, lambda a, b: mmr,
                                                      n_timesteps=500)
    assert n_it == 53 and hist.shape == (30, 106)
    np.testing.assert_allclose(Tf[[0, 10, 20, 29]],
                               [5116.149405501793, 2498.62652219369, 1716.1159109408936,
                                1235.4767222456146], rtol=1e-9)


def test_sharded_integrals_equal_global(default_case):
    """Wavelength shards with the global trapezoid weights reproduce np.trapz (multi-GPU contract)."""
    pl, P, T, lam, tabs = default_case
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    F = O.F_TOA(lam * 1e-4, pl['T_star'], a_rstar=pl['a_rstar'])
    full = O.emit(tabs, T, P, lam, F, pl['g'], pl['m_bar'], lambda a, b: mmr)
    wts = O.trapz_weights(lam * 1e-4)
    bol = np.zeros((30, 4))
    for lo, hi in ((0, 170), (170, 333), (333, 500)):
        t = {'1H2-16O': dict(P=tabs['1H2-16O']['P'], T=tabs['1H2-16O']['T'],
                             values=tabs['1H2-16O']['values'][:, :, lo:hi])}
        part = O.emit(t, T, P, lam[lo:hi], F[lo:hi], pl['g'], pl['m_bar'], lambda a, b: mmr,
                      trapz_w=wts[lo:hi])
        assert np.array_equal(part[0], full[0][:, lo:hi])
        bol += part[6]
    np.testing.assert_allclose(bol, full[6], rtol=1e-13)
    dT = O.thermo_from_bol(bol, T, P, pl['g'], pl['m_bar'], pl['alpha'], 'emit')
    np.testing.assert_allclose(dT, full[5], rtol=1e-9, atol=1e-9)


def test_golden_vectors(default_case):
    """Committed golden vectors (tests/golden/make_golden.py) still match the oracle."""
    path = os.path.join(GOLDEN, 'oracle_default_grid.json')
    with open(path) as fh:
        g = json.load(fh)
    pl, P, T, lam, tabs = default_case
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'], vmr=g['vmr'])
    spec, Tf, hist, dtaus, n_it = O.emission_spectrum(tabs, T, P, lam, pl, lambda a, b: mmr,
                                                      n_timesteps=1)
    idx = g['lam_index']
    np.testing.assert_allclose(spec[idx], g['spectrum'], rtol=1e-12)
    np.testing.assert_allclose(Tf, g['final_temps'], rtol=1e-12)
    np.testing.assert_allclose(dtaus[:, idx], g['dtaus'], rtol=1e-12)
    k, s = O.kappa(tabs, T[0], P[0], lam, mmr, pl['m_bar'])
    np.testing.assert_allclose(k[idx], g['kappa_layer0'], rtol=1e-13)
    np.testing.assert_allclose(s[idx], g['sigma'], rtol=1e-13)


# ---------------------------------------------------------------------------------------------
# Golden vectors produced by the reference's OWN source files (tests/golden/run_reference.py:
# frei/twostream.py, opacity.py, core.py, tp.py, chemistry.py executed under dependency stubs).
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def ref_run():
    with open(os.path.join(GOLDEN, 'reference_run.json')) as fh:
        return json.load(fh)


def test_reference_run_case_A_default_grid(ref_run, default_case):
    """Grid(T_ref=2400 K) + load_example_opacity(scale_factor=1) + emission_spectrum(1), mock chemistry."""
    a = ref_run['A']
    pl, P, T, lam, tabs = default_case
    assert abs(a['g_cgs'] - pl['g']) < 1e-9 * pl['g'] and abs(a['a_rstar'] - pl['a_rstar']) < 1e-12
    np.testing.assert_allclose(P, a['pressures_bar'], rtol=1e-14)
    np.testing.assert_allclose(T, a['init_temperatures'], rtol=1e-14)
    idx = a['lam_index']
    np.testing.assert_allclose(lam[idx], a['lam_um'], rtol=1e-14)
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    k, s = O.kappa(tabs, T[0], P[0], lam, mmr, pl['m_bar'])
    np.testing.assert_allclose(k[idx], a['kappa0'], rtol=1e-13)
    np.testing.assert_allclose(s[idx], a['sigma'], rtol=1e-12)
    spec, Tf, hist, dtaus, _ = O.emission_spectrum(tabs, T, P, lam, pl, lambda x, y: mmr, n_timesteps=1)
    # the reference's own fp64 noise in optically thin layers (DESIGN.md "conditioning") bounds this
    np.testing.assert_allclose(spec[idx], a['spectrum'], rtol=1e-9)
    np.testing.assert_allclose(Tf, a['final_temps'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(hist, a['temp_hist'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(dtaus[:, idx], a['dtaus'], rtol=1e-12)
    assert abs(O.effective_temperature(P, lam, spec, dtaus, Tf) - a['T_eff']) < 1e-6


def test_reference_run_case_B_full_solve(ref_run):
    """12 layers x 120 bins iterated to convergence: same number of history columns, same T."""
    b = ref_run['B']
    pl = O.hot_jupiter()
    P = O.pressure_grid(12, np.log10(1e-6), np.log10(200))
    T = O.temperature_grid(P, 2400.0, 0.1, 0.1)
    lam, _, _ = O.wavelength_grid(0.5, 10, 120)
    tabs = O.load_example_opacity(P, T, lam, scale_factor=1)
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    spec, Tf, hist, dtaus, n_it = O.emission_spectrum(tabs, T, P, lam, pl, lambda x, y: mmr,
                                                      n_timesteps=400)
    assert hist.shape[1] == b['n_columns'] == 2 * n_it
    np.testing.assert_allclose(Tf, b['final_temps'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(hist[:, -4:], b['temp_hist_last'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(spec, b['spectrum'], rtol=1e-7)
    np.testing.assert_allclose(dtaus[5], b['dtaus_row5'], rtol=1e-9)


def test_reference_run_case_C_layer_formulas(ref_run):
    """propagate_fluxes on both E branches and the scalar thermodynamics chain."""
    c = ref_run['C']
    lam_cm = np.array(c['lam_um']) * 1e-4
    F2u, F1d = O.propagate_fluxes(lam_cm, np.array(c['F1']), np.array(c['F2']), 1800.0, 1650.0,
                                  np.array(c['dtau']), np.array(c['w0']), 0)
    np.testing.assert_allclose(F2u, c['F2u'], rtol=1e-11)
    np.testing.assert_allclose(F1d, c['F1d'], rtol=1e-11)
    pl = O.hot_jupiter()
    p1, p2, T1, T2 = 1.0 * O.BAR, 0.6 * O.BAR, 1900.0, 1500.0
    for t in c['thermo']:
        Fb = (5e9 + t['dF'], 1e9, 5e9, 1e9)
        dT = O.layer_thermo(Fb, T1, T2, p1, p2, pl['g'], pl['m_bar'], pl['alpha'])
        assert abs(dT - t['dT']) <= 1e-11 * max(1.0, abs(t['dT']))
        assert abs(O.delta_z_i(T1, p1, p2, pl['g'], pl['m_bar']) - t['dz_cm']) < 1e-9 * t['dz_cm']


def test_reference_run_case_D_mock_chemistry(ref_run):
    d = ref_run['D']
    ref = O.mock_mmr(['1H2-16O', '12C-16O', '48Ti-16O'])
    for i, name in enumerate(['1H2-16O', '12C-16O', '48Ti-16O']):
        np.testing.assert_allclose(d['vmr'][name], 1.5e-3, rtol=1e-14)
        np.testing.assert_allclose(d['mmr'][name], ref[i], rtol=1e-14)


def test_diagnostics_restatements(default_case):
    """pressure_milne / contribution_function (frei/core.py:390-395, frei/plot.py:63-79): the
    Milne pressure is what effective_temperature_milne averages, the contribution function is
    normalised per wavelength and invariant under the wavelength-only and layer-constant factors."""
    rs = np.random.RandomState(7)
    L, n = 30, 64
    P = O.pressure_grid(L, np.log10(1e-6), np.log10(200))
    T = O.temperature_grid(P, 2400.0, 0.1, 0.1)
    lam = np.logspace(np.log10(0.5), 1, n)
    dtaus = np.vstack([np.ones(n), np.sort(10.0 ** rs.uniform(-6, 1, (L - 1, n)), axis=0)[::-1]])
    spec = rs.uniform(1e10, 1e13, n)
    pm = O.pressure_milne(P, dtaus)
    assert pm.shape == (n,) and np.all((pm >= P.min()) & (pm <= P.max()))
    t_m = np.interp(np.average(pm, weights=spec * lam * 1e-4), P[::-1], T[::-1])
    assert t_m == O.effective_temperature_milne(P, lam, spec, dtaus, T)
    cf = O.contribution_function(lam, P, T, dtaus)
    assert cf.shape == (L, n) and np.all(cf >= 0)
    np.testing.assert_allclose(cf.sum(axis=0), 1.0, rtol=1e-13)
    # level order: the top level (index L-1) has tau = its own dtau
    top = np.exp(-dtaus[-1]) * dtaus[-1] / np.expm1(O.h * O.c / O.k_B / (lam * 1e-4) / T[-1])
    bot_tau = dtaus.sum(axis=0)
    bot = np.exp(-bot_tau) * dtaus[0] / np.expm1(O.h * O.c / O.k_B / (lam * 1e-4) / T[0])
    np.testing.assert_allclose(cf[-1] / cf[0], top / bot, rtol=1e-9)


def test_reference_run_case_E_contribution_function(ref_run):
    """
    The array the reference's own frei/plot.py hands to pcolormesh (cf[::-1], plot.py:63-83),
    captured by running dashboard() from the reference checkout against a matplotlib stand-in
    (tests/golden/run_reference.py, case E): pins the oracle's contribution_function.  Inputs:
    the golden dtaus / final temperatures of case A on the same sampled wavelength columns (the
    function is column-wise independent).
    """
    a, e = ref_run['A'], ref_run['E']
    assert e['lam_index'] == a['lam_index']
    P = np.array(a['pressures_bar'])
    lam = np.array(a['lam_um'])
    cf = O.contribution_function(lam, P, np.array(a['final_temps']), np.array(a['dtaus']))
    ref = np.array(e['contribution_function'])
    assert cf.shape == ref.shape
    np.testing.assert_allclose(cf, ref, rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(cf.sum(axis=0), e['column_sums'], rtol=1e-13)
