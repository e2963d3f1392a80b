"""Sweep-kernel throughput vs wavelength count (and species/layers): evals/s from CUDA events."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from frei_b200 import synthetic  # noqa: E402
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F32, FREI_F64  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--L', type=int, default=50)
ap.add_argument('--S', type=int, default=3)
ap.add_argument('--table-dtype', type=int, default=64)
ap.add_argument('--flux-dtype', type=int, default=64)
ap.add_argument('--nlam', type=int, nargs='+', default=[50_000, 100_000, 151_552, 200_000, 303_104, 400_000, 800_000])
ap.add_argument('--plan', type=int, default=0, help='frei_b200_debug_plan: 0 auto, 1 32-wide, 2 64-wide, 3 mixed')
a = ap.parse_args()
from frei_b200 import _cabi  # noqa: E402
_cabi.check(_cabi.load().frei_b200_debug_plan(a.plan))
dt = FREI_F32 if a.table_dtype == 32 else FREI_F64
for n_lam in a.nlam:
    w = synthetic.make_workload(a.L, n_lam, a.S, 2400.0, table_f32=(dt == FREI_F32))
    tab = synthetic.device_table(w, dt)
    pl = w['planet']
    eng = Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                 alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                 flux_dtype=FREI_F32 if a.flux_dtype == 32 else FREI_F64)
    for _ in range(3):
        eng.sweep(FREI_EMIT); eng.sweep(FREI_ABSORB)
    eng.sweep_events = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        eng.sweep(FREI_EMIT); eng.sweep(FREI_ABSORB)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    sw = np.mean([x.elapsed_time(y) for x, y in eng.sweep_events])
    ev = (a.L - 1) * n_lam
    print(f'L {a.L} S {a.S} tab f{a.table_dtype} flux f{a.flux_dtype} n_lam {n_lam:8d}: step {ms:.3f} ms  sweep {sw:.4f} ms '
          f'-> kernel {ev / sw / 1e6:.1f} G evals/s, step {2 * ev / ms / 1e6:.1f} G evals/s')
    del eng, tab
    torch.cuda.empty_cache()
