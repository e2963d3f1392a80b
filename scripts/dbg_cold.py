"""Where does the cold out-of-table case differ from the 80-bit oracle? (debug helper)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import importlib.util
spec = importlib.util.spec_from_file_location('tg', os.path.join(os.path.dirname(__file__), '..', 'tests', 'test_gpu_parity.py'))
tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
from frei_b200 import synthetic
from frei_b200.engine import FREI_EMIT, FREI_ABSORB
LD = np.longdouble
L, n_lam, S, T_ref = 12, 129, 2, 120.0
w = synthetic.make_workload(L, n_lam, S, T_ref)
tabs = synthetic.host_tables(w)
ref = tg._oracle_iteration(w, tabs, 1)
refx = tg._oracle_iteration(w, tabs, 1, wd=LD)
eng = tg._engine(w)
np.set_printoptions(linewidth=200, precision=3)
for k, direction in enumerate((FREI_EMIT, FREI_ABSORB)):
    eng.sweep(direction, with_dtaus=True)
    Fu, Fd = eng.F_up[0].cpu().numpy(), eng.F_down[0].cpu().numpy()
    for name, g, x, r in (('Fu', Fu, refx[k]['Fu'], ref[k]['Fu']), ('Fd', Fd, refx[k]['Fd'], ref[k]['Fd'])):
        e = (np.abs(g.astype(LD) - x) / np.maximum(np.abs(x), LD(1e-250))).astype(np.float64)
        i, j = np.unravel_index(np.argmax(e), e.shape)
        print(k, name, 'max rel err vs ld', e.max(), 'at level', i, 'lam idx', j, 'gpu', g[i, j], 'ld', float(x[i, j]), 'f64', r[i, j])
        print('   per-level max:', e.max(axis=1))
    sums = eng.sums[0].cpu().numpy()
    print('sums gpu\n', sums, '\nld\n', np.asarray(refx[k]['bol'], dtype=np.float64))
    print('T', w['T_init'], 'dtaus range', ref[k]['dtaus'][1:].min(), ref[k]['dtaus'][1:].max())
