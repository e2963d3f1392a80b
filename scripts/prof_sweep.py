"""Short driver for ncu: a few RE iterations of BASELINE config C2 (or --config) on one GPU."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from frei_b200 import synthetic  # noqa: E402
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F32, FREI_F64  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--table-dtype', type=int, default=64)
ap.add_argument('--nlam', type=int, default=0)
ap.add_argument('--flux-dtype', type=int, default=64)
ap.add_argument('--plan', type=int, default=0)
a = ap.parse_args()
from frei_b200 import _cabi  # noqa: E402
_cabi.check(_cabi.load().frei_b200_debug_plan(a.plan))
L, n_lam, S, T_ref = synthetic.CONFIGS[a.config]
if a.nlam:
    n_lam = a.nlam
dt = FREI_F32 if a.table_dtype == 32 else FREI_F64
w = synthetic.make_workload(L, n_lam, S, T_ref, table_f32=(dt == FREI_F32))
tab = synthetic.device_table(w, dt)
pl = w['planet']
eng = Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
             alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'],
             flux_dtype=FREI_F32 if a.flux_dtype == 32 else FREI_F64)
for _ in range(a.iters):
    eng.sweep(FREI_EMIT)
    eng.sweep(FREI_ABSORB)
torch.cuda.synchronize()
print('T[0], T[-1] =', eng.T[0, 0].item(), eng.T[0, -1].item())
