import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import frei_b200 as frei
from frei_b200 import synthetic
from frei_b200.opacity import OpacityTable
for n_lam in (30001, 30000):
    w = synthetic.make_workload(40, n_lam, 3)
    tabs = synthetic.host_tables(w)
    op = {k: OpacityTable(t['values'], t['P'], t['T'], w['lam_um']) for k, t in tabs.items()}
    pl = w['planet']
    planet = frei.Planet(a_rstar=pl['a_rstar'], m_bar=pl['m_bar'], g=pl['g'] / 100.0, T_star=pl['T_star'], alpha=pl['alpha'])
    res = []
    for rep in range(3):
        grid = frei.Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
        grid.load_opacities(opacities=op)
        s, T, h, d = grid.emission_spectrum(n_timesteps=40)
        res.append((grid.n_iterations, T.copy(), np.asarray(s.flux).copy()))
        s2, T2, h2, d2 = grid.emission_spectrum(n_timesteps=40)       # cached engine, reset
        res.append((grid.n_iterations, T2.copy(), np.asarray(s2.flux).copy()))
    print(n_lam, [r[0] for r in res], [float(np.abs(r[1] - res[0][1]).max()) for r in res])
