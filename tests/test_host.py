"""Host-side mirror of the reference API: grids, names, chemistry mock, fixtures (no GPU)."""
import numpy as np
import pytest

import frei_b200 as frei
import importlib
from frei_b200 import units as U
from frei_b200.core import converged_layers, wavelength_grid
from frei_b200.sharding import shard_range, shard_ranges
from oracle import frei_oracle as O

chem = importlib.import_module('frei_b200.chemistry')   # the package attribute is the function


def test_grid_init():
    """frei/tests/test_core.py:8-16."""
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet=planet)
    for attr in ['lam', 'init_temperatures', 'pressures']:
        assert hasattr(grid, attr)
    assert len(U.value(grid.lam, 'um')) == 500 and len(U.value(grid.pressures, 'bar')) == 30
    assert 'Grid in T=' in repr(grid)
    with pytest.raises(ValueError, match='Must load opacities'):
        grid.emission_spectrum()


def test_grids_match_oracle():
    np.testing.assert_array_equal(U.value(frei.pressure_grid(30, -6, 1.1), 'bar'),
                                  O.pressure_grid(30, -6, 1.1))
    P = O.pressure_grid(40, -6, 2)
    np.testing.assert_array_equal(U.value(frei.temperature_grid(P, 2400, 0.1, 0.1), 'K'),
                                  O.temperature_grid(P, 2400.0, 0.1, 0.1))
    lam, bins, R = wavelength_grid(0.5, 10, 300)
    lam_o, bins_o, R_o = O.wavelength_grid(0.5, 10, 300)
    np.testing.assert_array_equal(U.value(lam, 'um'), lam_o)
    np.testing.assert_array_equal(bins, bins_o)
    assert R == R_o
    pl = frei.Planet.from_hot_jupiter()
    assert abs(U.gravity_cgs(pl.g) - O.hot_jupiter()['g']) < 1e-9
    assert abs(pl.a_rstar - O.hot_jupiter()['a_rstar']) < 1e-12


def test_example_opacity_matches_reference_fixture_restated_in_oracle():
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet, T_ref=2400)
    op = frei.load_example_opacity(grid, scale_factor=1)
    assert "1H2-16O" in op
    tab = op["1H2-16O"]
    for attr in ['wavelength', 'temperature', 'pressure']:
        assert hasattr(tab, attr)
    ref = O.load_example_opacity(U.value(grid.pressures, 'bar'), U.value(grid.init_temperatures, 'K'),
                                 U.value(grid.lam, 'um'), scale_factor=1)["1H2-16O"]
    vals = np.asarray(tab.values)
    np.testing.assert_array_equal(vals[0, 0], ref['values'][0, 0])
    assert abs(vals[0, 0, 0] - 40.01236740950614) < 1e-12          # SURVEY appendix B anchor
    from frei_b200.engine import normalise_table
    Pax, Tax, v, has_T = normalise_table(tab)
    np.testing.assert_array_equal(Pax, ref['P'])
    np.testing.assert_array_equal(Tax, ref['T'])
    np.testing.assert_array_equal(v, ref['values'])
    assert has_T


@pytest.mark.parametrize("iso, species", list(zip(['1H2-16O', 'Na', 'K', '48Ti-16O'],
                                              ["H2O", "Na", "K", "TiO"])))
def test_chemical_names_manipulation_0(iso, species):
    assert chem.iso_to_species(iso) == species


@pytest.mark.parametrize("species, fastchem", list(zip(
    ['H2O', 'TiO', 'VO', 'Na', 'K', 'CO', 'CrH', 'CF4O', 'Al2Cl6', 'AlNaF4', 'ClAlF2'],
    ['H2O1', 'O1Ti1', 'O1V1', 'Na', 'K', 'C1O1', 'Cr1H1', 'C1F4O1', 'Al2Cl6', 'Al1F4Na1',
     'Al1Cl1F2'])))
def test_chemical_names_manipulation_1(species, fastchem):
    assert chem.species_name_to_fastchem_name(species) == fastchem


@pytest.mark.parametrize("species, iso", list(zip(
    ['H2O', 'TiO', 'VO', 'Na', 'K', 'CO', 'CrH', 'CF4O', 'Al2Cl6', 'AlClF2'],
    ['1H2-16O', '48Ti-16O', '51V-16O', 'Na', 'K', '12C-16O', '52Cr-1H', '12C-19F4-16O',
     '27Al2-35Cl6', '27Al-35Cl-19F2'])))
def test_chemical_names_manipulation_2_and_3(species, iso):
    assert chem.species_name_to_common_isotopologue_name(species) == iso
    assert chem.species_name_to_common_isotopologue_name(chem.iso_to_species(iso)) == iso


def test_mock_chemistry_matches_reference_mock():
    """No pyfastchem -> VMR 1.5e-3 for every species, mmr = vmr * mass / m_bar (chemistry.py:197-246)."""
    T = np.array([1000.0, 2000.0, 3000.0])
    P = np.array([10.0, 1.0, 0.1])
    species = ['1H2-16O', '12C-16O', '12C-1H4']
    mmr, vmr = frei.chemistry(T, P, species, return_vmr=True)
    ref = O.mock_mmr(species)
    for i, s in enumerate(species):
        np.testing.assert_allclose(vmr[s], 1.5e-3, rtol=1e-15)
        np.testing.assert_allclose(mmr[s], ref[i], rtol=1e-15)
    assert [O.iso_to_mass(s) for s in species] == [18.0, 28.0, 16.0]


def test_convergence_rule_matches_oracle():
    rs = np.random.RandomState(2)
    hists = [rs.uniform(900, 1100, (7, 2)) for _ in range(6)]
    dT = rs.uniform(-6, 6, 7)
    a, h1 = converged_layers(hists, dT, 2, 3.0)
    b, h2 = O.converged_layers(hists, dT, 2, 3.0)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(h1, h2)


def test_shard_ranges_partition_the_axis():
    for n, w in [(10, 3), (200000, 8), (7, 8), (1000001, 4)]:
        r = shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
        assert r[1] == shard_range(n, 1, w)


def test_trapezoid_weights_reproduce_np_trapz():
    from frei_b200.core import _trapz_weights_cm
    rs = np.random.RandomState(3)
    for n in (1, 2, 3, 500):
        lam = np.sort(rs.uniform(0.5, 10, n))
        F = rs.uniform(1, 2, n)
        trapz = getattr(np, 'trapezoid', None) or np.trapz
        np.testing.assert_allclose((_trapz_weights_cm(lam) * F).sum(), trapz(F, lam * 1e-4), rtol=1e-14, atol=0)


def test_emission_spectrum_argument_errors():
    """Same error behaviour as frei/core.py:259-260; the gather mode is validated before any GPU work."""
    import pytest
    grid = frei.Grid(frei.Planet.from_hot_jupiter())
    with pytest.raises(ValueError, match='Must load opacities'):
        grid.emission_spectrum()
    grid.opacities = {'1H2-16O': None}
    with pytest.raises(ValueError, match='gather'):
        grid.emission_spectrum(gather='rank0')
    with pytest.raises(ValueError, match='emission_spectrum'):
        grid.diagnostics()


def test_bench_helpers_without_a_gpu():
    """bench.py's host-side pieces: committed ncu traffic entry, peak fallback, clock-sample parsing."""
    import bench
    traffic, src = bench.read_traffic('C2', 64, 64, 1)
    assert traffic and 3e8 < traffic < 6e8 and 'ncu' in src
    assert bench.read_traffic('C9', 64, 64, 1) == (None, None)
    peak, how = bench.read_peaks()
    assert 5000 < peak < 9000 and how
    cs = bench.ClockSampler(0)
    cs.proc = object.__new__(type('P', (), {'terminate': lambda s: None, 'wait': lambda s, timeout=None: 0,
                                            'kill': lambda s: None}))
    cs.lines = ['0, 1965, 1965, 640.2, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active',
                '0, 1800, 1965, 990.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active',
                'garbage']
    out = cs.stop()
    assert out['sm_mhz'] == 1882.5 and out['sm_max_mhz'] == 1965.0
    assert out['reasons'] == ['sw_power_cap'] and out['samples'] == 2 and out['power_w_max'] == 990.0
    assert set(bench.WORKLOADS) == {'C1', 'C2', 'C3', 'C5'}
