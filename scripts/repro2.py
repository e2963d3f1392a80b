import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frei_b200 import synthetic
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F64
for L, n_lam, S in ((30, 30000, 3), (40, 30000, 3), (40, 3000, 3), (40, 30000, 1), (50, 30000, 3), (40, 512, 3)):
    w = synthetic.make_workload(L, n_lam, S)
    tab = synthetic.device_table(w, FREI_F64)
    pl = w['planet']
    outs = []
    for rep in range(4):
        eng = Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                     alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'])
        for _ in range(6):
            eng.iteration()
        torch.cuda.synchronize()
        outs.append((eng.T.cpu().numpy().copy(), eng.sums.cpu().numpy().copy()))
    print(L, n_lam, S, 'T maxdiff over reps', [float(np.abs(o[0] - outs[0][0]).max()) for o in outs],
          'sums rel', [float(np.abs(o[1] - outs[0][1]).max() / np.abs(outs[0][1]).max()) for o in outs])
