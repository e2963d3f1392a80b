"""Where does the end-to-end time of Grid.emission_spectrum go? (C2, one GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frei_b200 import synthetic
from frei_b200.core import Grid, Planet
from frei_b200.engine import FREI_F64, FREI_EMIT, FREI_ABSORB
L, n_lam, S, T_ref = synthetic.CONFIGS['C2']
w = synthetic.make_workload(L, n_lam, S, T_ref)
table = synthetic.device_table(w, FREI_F64)
pl = w['planet']
planet = Planet(a_rstar=pl['a_rstar'], m_bar=pl['m_bar'], g=pl['g'] / 100.0, T_star=pl['T_star'], alpha=pl['alpha'])
grid = Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
grid.attach_device_table(table, species=w['species'])
K = 20
grid.emission_spectrum(n_timesteps=2, n_zero_crossings=10 ** 9, convergence_dT=0)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    out = grid.emission_spectrum(n_timesteps=K, n_zero_crossings=10 ** 9, convergence_dT=0)
    torch.cuda.synchronize()
    print('emission_spectrum(K=%d): %.3f ms' % (K, 1e3 * (time.perf_counter() - t0)))
    del out
from frei_b200 import core
core.TRACE = []
out = grid.emission_spectrum(n_timesteps=K, n_zero_crossings=10 ** 9, convergence_dT=0)
marks, core.TRACE = core.TRACE, None
del out
print('marks inside emission_spectrum [ms since enter]: ' + ' | '.join('%s %.3f' % (lab, 1e3 * (t - marks[0][1])) for lab, t in marks))
eng = grid.engine
# phases
torch.cuda.synchronize(); t0 = time.perf_counter()
eng.reset(w['T_init']); torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(K):
    eng.iteration(); eng.read_history()
t2 = time.perf_counter()
for _ in range(K):
    eng.iteration()
torch.cuda.synchronize(); t3 = time.perf_counter()
eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True); torch.cuda.synchronize(); t4 = time.perf_counter()
buf = torch.empty((L + 1, n_lam), dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t5 = time.perf_counter()
buf[1:].copy_(eng.dtaus[0], non_blocking=True); torch.cuda.synchronize(); t6 = time.perf_counter()
buf[1:].copy_(eng.dtaus[0], non_blocking=True); torch.cuda.synchronize(); t7 = time.perf_counter()
print('reset %.3f ms | %d x (iteration + read_history) %.3f ms | %d x iteration (no sync) %.3f ms | final emit %.3f ms | pin 80 MB %.1f ms | D2H 80 MB %.3f / %.3f ms (%.1f GB/s)'
      % (1e3 * (t1 - t0), K, 1e3 * (t2 - t1), K, 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5), 1e3 * (t7 - t6), 0.08 / (t7 - t6)))
