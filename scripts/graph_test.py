"""CUDA-graph replay of the RE iteration: same results as individual launches, step time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frei_b200 import synthetic
from frei_b200.engine import Engine, FREI_F64

w = synthetic.make_workload(50, 200000, 3)
tab = synthetic.device_table(w, FREI_F64)
pl = w['planet']
def mk():
    return Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                  alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'])
a, b = mk(), mk()
print('captured', b.capture_iteration())
for _ in range(5):
    a.iteration(); b.iteration()
torch.cuda.synchronize()
print('T equal', torch.equal(a.T, b.T), 'F_up equal', torch.equal(a.F_up, b.F_up), 'F_down equal', torch.equal(a.F_down, b.F_down))
for eng, name in ((a, 'launches'), (b, 'graph')):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): eng.iteration()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f'{name}: {ms:.4f} ms/step -> {2 * 49 * 200000 / ms / 1e6:.2f} G evals/s')
