"""Timeline of one sweep -> post -> sweep hand-over from globaltimer stamps (probe build -DPOST_STAMPS only)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frei_b200 import synthetic, _cabi
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F64
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
L, n_lam, S, T_ref = synthetic.CONFIGS[cfg]
if len(sys.argv) > 2: n_lam = int(sys.argv[2])
w = synthetic.make_workload(L, n_lam, S, T_ref)
tab = synthetic.device_table(w, FREI_F64)
pl = w['planet']
eng = Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'], alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'])
lib = eng.lib
lib.frei_b200_debug_stamps.restype = C.c_int
lib.frei_b200_debug_stamps.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
names = {28: 'sweep A: first warp done', 29: 'sweep A: last warp done', 0: 'post: first CTA entry', 1: 'post: last CTA entry',
         2: 'post: first CTA released', 3: 'post: last CTA released', 4: 'post: stage 1 first', 5: 'post: stage 1 last',
         7: 'post: last ticket', 9: 'post: stage 2', 11: 'post: T update (stores, tracker)', 23: 'post:   thread 0 past dT', 25: 'post:   thread 1 dT computed', 27: 'post:   all dT computed', 31: 'post:   first bracket written', 15: 'post:   all brackets written', 13: 'post: records written',
         14: 'sweep B: first CTA resident', 16: 'sweep B: first CTA released', 17: 'sweep B: last CTA released',
         18: 'sweep B: records in smem first', 19: 'sweep B: records in smem last', 20: 'sweep B: first warp done', 21: 'sweep B: last warp done'}
flux = eng._flux_struct(False)
st = eng._stream()
for _ in range(5):
    eng.sweep(FREI_EMIT); eng.sweep(FREI_ABSORB)
torch.cuda.synchronize()
acc = []
for rep in range(6):
    # sequence: [emit sweep][post][absorb sweep] ; stamps of the emit sweep's end come from sweep (20, 21) of BOTH sweeps, so isolate:
    eng.sweep(FREI_EMIT); torch.cuda.synchronize()
    lib.frei_b200_debug_stamps(None, 1)
    eng.sweep(FREI_ABSORB)                                   # sweep A (absorb) + post A
    _cabi.check(lib.frei_b200_sweep(C.byref(eng._tab), C.byref(eng._spec), C.byref(eng._atm), C.byref(flux), FREI_EMIT,
                                    C.byref(eng._ws), st))    # sweep B (emit) alone
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)()
    lib.frei_b200_debug_stamps(out, 0)
    acc.append(np.array(list(out), dtype=np.float64))
a = np.array(acc)
# reference: last warp of sweep A done = min over... stamps 20/21 hold A and B; use 'post released first' (2) as origin instead
print(cfg, n_lam, 'origin = last warp of sweep A done; medians over 6 repetitions, microseconds')
for i in [28, 29, 0, 1, 2, 3, 4, 5, 7, 9, 23, 25, 27, 11, 31, 15, 13, 14, 16, 17, 18, 19, 20, 21]:
    print(f'  {names[i]:34s} {np.median((a[:, i] - a[:, 29])) / 1e3:9.2f}')
