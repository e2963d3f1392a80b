"""C4 batch on one GPU: which atmospheres of the T_eq x log g x metallicity grid do not meet
Grid.emission_spectrum's convergence rule within the iteration cap, and when do they stop?"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from frei_b200 import synthetic  # noqa: E402
from frei_b200.engine import Engine, FREI_F64  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--side', type=int, default=16)
ap.add_argument('--nlam', type=int, default=20_000)
ap.add_argument('--cap', type=int, default=3000)
a = ap.parse_args()
L, S = 50, 3
w = synthetic.make_workload(L, a.nlam, S, 2400.0)
pl = w['planet']
T_ref = np.linspace(1000, 2500, a.side)
logg = np.linspace(2.5, 4.0, a.side)
met = np.linspace(-1, 2, a.side)
tt, gg, mm = [x.ravel() for x in np.meshgrid(T_ref, logg, met, indexing='ij')]
B = tt.size
T0 = tt[:, None] * (w['P_bar'][None, :] / 0.1) ** 0.1
mmr = w['mmr'][None] * (10.0 ** mm)[:, None, None]
table = synthetic.device_table(w, FREI_F64)
eng = Engine(table, w['lam_um'], np.broadcast_to(w['P_bar'], (B, L)), T0, mmr, g=10.0 ** gg,
             m_bar=pl['m_bar'], alpha=1.0, T_star=pl['T_star'], a_rstar=pl['a_rstar'],
             ftoa_scale=(tt / 2400.0) ** 4)
iters, T = eng.solve_batch(a.cap, check_every=8)
torch.cuda.synchronize()
hist = np.bincount(np.minimum(iters // 50, 80))
print('iterations histogram (bins of 50):', {int(50 * i): int(c) for i, c in enumerate(hist) if c})
for cap in (400, 1000, 2000, a.cap):
    print(f'not converged within {cap}: {(iters >= cap).sum()} of {B}')
slow = np.argsort(-iters)[:12]
print('slowest:', [(float(tt[i]), round(float(gg[i]), 2), round(float(mm[i]), 2), int(iters[i])) for i in slow])
late = iters >= 400
if late.any():
    print('>= 400 iterations: T_ref range', tt[late].min(), tt[late].max(), 'log g', gg[late].min(), gg[late].max(),
          'met', mm[late].min(), mm[late].max())
json.dump({'T_ref': tt.tolist(), 'logg': gg.tolist(), 'met': mm.tolist(), 'iters': iters.tolist()},
          open(os.path.join('gpurun_out', 'r02_c4_iters.json'), 'w'))
