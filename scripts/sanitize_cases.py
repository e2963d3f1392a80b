"""CI-sized pass over every kernel of libfrei_b200.so, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_cases.py
    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py
Covers the sweep (32-wide, 64-wide and mixed plans; S = 1, 3, 5, 8; odd wavelength counts; dtaus),
post_kernel (fused reduction + update + re-bracketing, batch tracker), the split reduce / update
sequence, the fp32 sweep, kappa, propagate, diagnostics, binning (unit and x-spaced) and regrid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
import frei_b200 as frei  # noqa: E402
from frei_b200 import synthetic, _cabi  # noqa: E402
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F32, FREI_F64  # noqa: E402

lib = _cabi.load()


def engine(L, n_lam, S, B=1, flux=FREI_F64, tab=FREI_F64):
    w = synthetic.make_workload(L, n_lam, S, table_f32=(tab == FREI_F32))
    table = synthetic.device_table(w, tab)
    pl = w['planet']
    T0 = np.broadcast_to(w['T_init'], (B, L)) * np.linspace(0.8, 1.1, B)[:, None]
    return w, Engine(table, w['lam_um'], np.broadcast_to(w['P_bar'], (B, L)), T0,
                     np.broadcast_to(w['mmr'], (B, L, S)), g=pl['g'], m_bar=pl['m_bar'], alpha=pl['alpha'],
                     T_star=pl['T_star'], a_rstar=pl['a_rstar'], flux_dtype=flux)


def iterate(eng, n=2):
    for _ in range(n):
        eng.sweep(FREI_EMIT)
        eng.sweep(FREI_ABSORB)
    eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True)
    torch.cuda.synchronize()
    assert torch.isfinite(eng.T).all()


for plan in (0, 1, 2, 3):
    _cabi.check(lib.frei_b200_debug_plan(plan))
    for (L, n_lam, S) in ((12, 1000, 3), (20, 514, 8), (9, 333, 1), (7, 260, 5)):
        iterate(engine(L, n_lam, S)[1])
    print('sweep plan', plan, 'ok', flush=True)
_cabi.check(lib.frei_b200_debug_plan(0))

w, eng = engine(10, 600, 3, B=3)
iters, T = eng.solve_batch(6, check_every=2)
print('batch ok', iters, flush=True)

w, eng = engine(10, 600, 3)                       # split sequence: reduce, update_T
eng.layer_prep()
flux = eng._flux_struct(False)
st = eng._stream()
_cabi.check(lib.frei_b200_sweep(C.byref(eng._tab), C.byref(eng._spec), C.byref(eng._atm), C.byref(flux), FREI_EMIT,
                                C.byref(eng._ws), st))
_cabi.check(lib.frei_b200_reduce(C.byref(eng._atm), C.byref(eng._ws), eng.n_lam, st))
_cabi.check(lib.frei_b200_update_T(C.byref(eng._tab), C.byref(eng._atm), C.byref(eng._ws), FREI_EMIT, -1.0, None, st))
torch.cuda.synchronize()
print('split ok', flush=True)

for n_lam in (1024, 333, 514):
    iterate(engine(10, n_lam, 3, flux=FREI_F32, tab=FREI_F32)[1])
print('fp32 ok', flush=True)

w, eng = engine(10, 500, 3)
k, sg = eng.kappa()
planet = frei.Planet.from_hot_jupiter()
grid = frei.Grid(planet, n_wl_bins=200, n_layers=8, T_ref=1800)
grid.load_opacities(opacities=frei.load_example_opacity(grid, scale_factor=1))
spec, T, hist, dtaus = grid.emission_spectrum(n_timesteps=3)
d = grid.diagnostics(contribution_function=True, pressure_milne=True)
lam = np.logspace(-0.3, 1, 64)
F = np.ones(64)
frei.propagate_fluxes(lam, F, F, 1500.0, 1400.0, np.full(64, 0.3), omega_0=np.full(64, 0.2))
torch.cuda.synchronize()
print('api ok', float(d['T_eff']), flush=True)

from frei_b200.interp import groupby_bins_agg  # noqa: E402
from frei_b200.opacity import bin_and_regrid  # noqa: E402
rs = np.random.RandomState(0)
wl = np.sort(rs.uniform(0.46, 9.9, 5000))
a = rs.uniform(0, 1, (2, 3, 5000))
edges = np.concatenate([[0.45], np.logspace(np.log10(0.5), 1, 100)])
groupby_bins_agg(a, wl, edges)
groupby_bins_agg(a.astype(np.float32), wl, edges)
for groupies in (True, False):
    bin_and_regrid(a, wl, [500.0, 900.0], [0.1, 1.0, 10.0], [600.0, 700.0, 1000.0], [0.5, 5.0], edges,
                   lam=0.5 * (edges[1:] + edges[:-1]), groupies=groupies)
torch.cuda.synchronize()
print('binning ok', flush=True)
print('ALL CASES DONE')
