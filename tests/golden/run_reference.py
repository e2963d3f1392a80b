"""
Executes the reference's OWN source files (frei/twostream.py, opacity.py, core.py, tp.py,
chemistry.py under /root/reference) with the dependency stubs of tests/golden/refstubs
(astropy.units/constants, xarray, specutils, periodictable — none installable here) and
writes golden vectors to tests/golden/reference_run.json.  frei/phoenix.py and interp.py are not
on the path and are replaced by empty modules; frei/plot.py runs against a matplotlib stand-in that
hands back the array it is asked to draw (the contribution function, frei/plot.py:63-83).

The stubs are ours, so what this pins is the reference's arithmetic and control flow
(sweep order, stale reads, top pseudo-layer, thermodynamics incl. astropy's unit algebra,
convergence rule, final emit) — not scipy's interpolation, which the stub calls directly.

    python tests/golden/run_reference.py            # only works where /root/reference exists
"""
import importlib
os_environ_set = __import__('os').environ.setdefault('TQDM_DISABLE', '1')
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FREI_REFERENCE', '/root/reference')


def load_reference():
    sys.path.insert(0, os.path.join(HERE, 'refstubs'))
    pkg = types.ModuleType('frei')
    pkg.__path__ = [os.path.join(REF, 'frei')]
    sys.modules['frei'] = pkg
    for name, attrs in (('frei.phoenix', ['get_binned_phoenix_spectrum']),
                        ('frei.interp', ['groupby_bins_agg'])):
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, lambda *x, **k: (_ for _ in ()).throw(NotImplementedError(a)))
        sys.modules[name] = mod
    mods = {m: importlib.import_module(f'frei.{m}')
            for m in ('tp', 'chemistry', 'opacity', 'twostream', 'plot', 'core')}
    return mods


def main():
    import warnings
    warnings.simplefilter('ignore')
    m = load_reference()
    import astropy.units as u
    core, opacity, twostream, chemistry = m['core'], m['opacity'], m['twostream'], m['chemistry']
    out = {}

    # ---- case A: the reference's own test (frei/tests/test_core.py:19-71), mock chemistry ----
    planet = core.Planet.from_hot_jupiter()
    grid = core.Grid(planet=planet, T_ref=2400 * u.K)
    op = grid.load_opacities(opacities=opacity.load_example_opacity(grid, scale_factor=1))
    k, sig = opacity.kappa(op, grid.init_temperatures[0], grid.pressures[0], grid.lam, m_bar=planet.m_bar)
    flux_unit = u.erg / u.s / u.cm ** 3
    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=1)
    idx = list(range(0, 500, 7))
    out['A'] = dict(
        g_cgs=float(planet.g.to(u.cm / u.s ** 2).value), a_rstar=float(planet.a_rstar),
        pressures_bar=grid.pressures.to(u.bar).value.tolist(),
        init_temperatures=grid.init_temperatures.to(u.K).value.tolist(),
        lam_um=grid.lam.to(u.um).value[idx].tolist(), lam_index=idx,
        kappa0=k.to(u.cm ** 2 / u.g).value[idx].tolist(),
        sigma=sig.to(u.cm ** 2 / u.g).value[idx].tolist(),
        spectrum=spec.flux.to(flux_unit).value[idx].tolist(),
        final_temps=temps.to(u.K).value.tolist(), temp_hist=hist.to(u.K).value.tolist(),
        dtaus=np.asarray(dtaus)[:, idx].tolist(),
        T_eff=float(core.effective_temperature(grid, spec, dtaus, temps).to(u.K).value))

    # ---- case E: the contribution function the reference's dashboard draws for case A ----
    # (frei/plot.py:63-83 executed from the reference's own file; the matplotlib stand-in raises
    # Captured with the arguments of ax[1].pcolormesh(lg, pg, cf[::-1], ...))
    import matplotlib
    try:
        m['plot'].dashboard(grid.lam, spec.flux, 0 * spec.flux, dtaus, grid.pressures, temps, hist, op)
        raise RuntimeError('dashboard did not reach pcolormesh')
    except matplotlib.Captured as cap:
        cf = np.asarray(cap.payload[2])
    out['E'] = dict(lam_index=idx, contribution_function=cf[:, idx].tolist(),
                    column_sums=cf.sum(axis=0)[idx].tolist())

    # ---- case B: full solve on a small grid (convergence rule, many iterations) ----
    grid = core.Grid(planet=planet, T_ref=2400 * u.K, n_layers=12, n_wl_bins=120)
    grid.load_opacities(opacities=opacity.load_example_opacity(grid, scale_factor=1))
    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=400)
    out['B'] = dict(n_layers=12, n_wl_bins=120, n_columns=int(hist.shape[1]),
                    spectrum=spec.flux.to(flux_unit).value.tolist(),
                    final_temps=temps.to(u.K).value.tolist(),
                    temp_hist_last=hist.to(u.K).value[:, -4:].tolist(),
                    dtaus_row5=np.asarray(dtaus)[5].tolist())

    # ---- case C: propagate_fluxes on both E branches, and the layer thermodynamics helpers ----
    rs = np.random.RandomState(5)
    n = 64
    lam = np.logspace(np.log10(0.5), 1, n) * u.um
    F1 = 10 ** rs.uniform(9, 14, n) * flux_unit
    F2 = 10 ** rs.uniform(9, 14, n) * flux_unit
    dtau = 10 ** rs.uniform(-4, 2, n)
    w0 = np.concatenate([10 ** rs.uniform(-8, -1.1, n // 2), rs.uniform(0.11, 0.9, n // 2)])
    F2u, F1d = twostream.propagate_fluxes(lam, F1, F2, 1800 * u.K, 1650 * u.K, dtau, omega_0=w0, g_0=0)
    g = planet.g
    p1, p2, T1, T2 = 1.0 * u.bar, 0.6 * u.bar, 1900 * u.K, 1500 * u.K
    thermo = []
    for dF in (3.0e6, -2.5e7, 0.0):
        bol = [(5e9 + dF) * flux_unit * u.cm, 1e9 * flux_unit * u.cm, 5e9 * flux_unit * u.cm,
               1e9 * flux_unit * u.cm]
        div, dz = twostream.div_bol_net_flux(bol[0], bol[1], bol[2], bol[3], T1, T2, p1, p2, g,
                                             alpha=planet.alpha, m_bar=planet.m_bar)
        dt = twostream.delta_t_i(p1, p2, T1, T2, div, g, m_bar=planet.m_bar)
        dT = twostream.delta_temperature(div, p1, p2, T1, dt, g).decompose()
        thermo.append(dict(dF=dF, dT=float(dT.to(u.K).value), dz_cm=float(dz.to(u.cm).value)))
    out['C'] = dict(lam_um=lam.value.tolist(), F1=F1.value.tolist(), F2=F2.value.tolist(),
                    dtau=dtau.tolist(), w0=w0.tolist(),
                    F2u=F2u.to(flux_unit).value.tolist(), F1d=F1d.to(flux_unit).value.tolist(),
                    thermo=thermo)

    # ---- case D: mock chemistry ----
    mmr, vmr = chemistry.chemistry(np.array([1000.0, 2000.0]) * u.K, np.array([1.0, 0.1]) * u.bar,
                                   ['1H2-16O', '12C-16O', '48Ti-16O'], return_vmr=True, m_bar=planet.m_bar)
    out['D'] = dict(mmr={k_: v.tolist() for k_, v in mmr.items()}, vmr={k_: v.tolist() for k_, v in vmr.items()})

    with open(os.path.join(HERE, 'reference_run.json'), 'w') as fh:
        json.dump(out, fh)
    a = out['A']
    print('case A: peak', max(a['spectrum']), 'T_eff', a['T_eff'], '| case B columns', out['B']['n_columns'])


if __name__ == '__main__':
    main()
