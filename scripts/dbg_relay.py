"""Where does the relay plan differ from the plan of whole chunks?  (debug aid for the bit-identity test)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frei_b200 import synthetic, _cabi
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F64
L, n_lam, S = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (30, 160000, 3)))
force = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = _cabi.load()
w = synthetic.make_workload(L, n_lam, S)
tab = synthetic.device_table(w, FREI_F64)
pl = w['planet']
out = {}
for plan in (2, force):
    _cabi.check(lib.frei_b200_debug_plan(plan))
    eng = Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                 alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'])
    snaps = []
    for it in range(2):
        for d in (FREI_EMIT, FREI_ABSORB):
            eng.sweep(d)
            torch.cuda.synchronize()
            snaps.append((f'it{it} dir{d}', eng.F_up[0].clone(), eng.F_down[0].clone(), eng.sums[0].clone(), eng.T[0].clone()))
    out[plan] = snaps
_cabi.check(lib.frei_b200_debug_plan(0))
NS = L - 1
for (name, *a), (_, *b) in zip(out[2], out[force]):
    for nm, x, y in zip(('F_up', 'F_down', 'sums', 'T'), a, b):
        if torch.equal(x, y):
            continue
        d = (x != y)
        idx = d.nonzero()
        print(f'{name} {nm}: {int(d.sum())} of {d.numel()} differ; first {idx[0].tolist()} last {idx[-1].tolist()}')
        if nm in ('F_up', 'F_down'):
            lev = idx[:, 0].cpu().numpy(); lam = idx[:, 1].cpu().numpy()
            ch = lam // 64
            print('   levels', np.unique(lev)[:20], ' chunks', np.unique(ch)[:12], '... n chunks', len(np.unique(ch)))
            rel = ((x - y).abs() / y.abs().clamp_min(1e-300))[d].max().item()
            print('   max rel diff', rel)
    if not all(torch.equal(x, y) for x, y in zip(a, b)):
        break
else:
    print('bit-identical')
