"""cProfile of Grid.emission_spectrum(K=20) on C2 (host-side cost of the end-to-end call)."""
import cProfile, pstats, os, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frei_b200 import synthetic
from frei_b200.core import Grid, Planet
from frei_b200.engine import FREI_F64
L, n_lam, S, T_ref = synthetic.CONFIGS['C2']
w = synthetic.make_workload(L, n_lam, S, T_ref)
table = synthetic.device_table(w, FREI_F64)
pl = w['planet']
planet = Planet(a_rstar=pl['a_rstar'], m_bar=pl['m_bar'], g=pl['g'] / 100.0, T_star=pl['T_star'], alpha=pl['alpha'])
grid = Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
grid.attach_device_table(table, species=w['species'])
K = 20
for _ in range(3):
    grid.emission_spectrum(n_timesteps=K, n_zero_crossings=10 ** 9, convergence_dT=0)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(5):
    out = grid.emission_spectrum(n_timesteps=K, n_zero_crossings=10 ** 9, convergence_dT=0)
pr.disable()
print('mean per call under cProfile: %.3f ms' % (1e3 * (time.perf_counter() - t0) / 5))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
print(s.getvalue()[:6000])
