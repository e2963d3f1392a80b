"""Where the time of the kernel that follows every sweep goes: back-to-back timings (CUDA events) of
post (reduce + update + K0), reduce only, update + K0 only, K0 only, and of the sweep with and
without its post kernel.  usage: python scripts/post_probe.py [n_lam ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from frei_b200 import synthetic, _cabi  # noqa: E402
from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F64  # noqa: E402


def timed(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


import argparse  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument('nlam', type=int, nargs='*', default=[5_000, 200_000, 800_000])
ap.add_argument('--L', type=int, default=50)
ap.add_argument('--S', type=int, default=3)
args = ap.parse_args()
for n_lam in args.nlam:
    w = synthetic.make_workload(args.L, n_lam, args.S, 2400.0)
    tab = synthetic.device_table(w, FREI_F64)
    pl = w['planet']
    eng = Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                 alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'])
    for _ in range(3):
        eng.sweep(FREI_EMIT); eng.sweep(FREI_ABSORB)
    lib, st = eng.lib, eng._stream()
    T0 = eng.T.clone()
    a = (C.byref(eng._tab), C.byref(eng._atm), C.byref(eng._ws))
    flux = eng._flux_struct(False)

    def post():
        _cabi.check(lib.frei_b200_post(*a, eng.n_lam, FREI_EMIT, -1.0, None, 1, st))

    def reduce_():
        _cabi.check(lib.frei_b200_reduce(a[1], a[2], eng.n_lam, st))

    def update():
        _cabi.check(lib.frei_b200_update_T(*a, FREI_EMIT, -1.0, None, st))

    def prep():
        _cabi.check(lib.frei_b200_layer_prep(*a, None, None, None, None, None, st))

    def sweep_only():
        _cabi.check(lib.frei_b200_sweep(C.byref(eng._tab), C.byref(eng._spec), C.byref(eng._atm),
                                        C.byref(flux), FREI_EMIT, C.byref(eng._ws), st))

    def sweep_post():
        sweep_only(); post()

    out = {}
    for name, fn in (('K0', prep), ('reduce', reduce_), ('update+K0', update), ('post', post),
                     ('sweep', sweep_only), ('sweep+post', sweep_post)):
        eng.T.copy_(T0)
        out[name] = timed(fn, 100 if 'sweep' in name else 200)
    print(f'n_lam {n_lam}: ' + '  '.join(f'{k} {v:.1f} us' for k, v in out.items()) +
          f'  -> post in situ {out["sweep+post"] - out["sweep"]:.1f} us')
