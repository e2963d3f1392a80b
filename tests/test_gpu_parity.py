"""
GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on identical seeded inputs.  Tolerances (BASELINE.json north_star):
bit-exact bracket indices / out-of-bounds mask; per-wavelength fluxes within
1e-6 relative in fp64; converged T-P profiles within 0.1 K.

Flux comparisons use three references ("conditioning"): the fp64
oracle (= the reference's arithmetic, whose B'/(2E) (chi - psi - xi) grouping
loses up to ~4e-6 relative accuracy where delta_tau < 1e-6), the same formulas
in 80-bit extended precision (np.longdouble, noise ~5e-9 in the same places) and,
on selected wavelength columns, in 40-digit arithmetic (mpmath).  The kernel
must match the 40-digit value to RTOL_EXACT, the 80-bit value to RTOL_X
everywhere, and the fp64 oracle to 1e-6 except where the fp64 oracle itself is
further than that from its own exact value — there the kernel must be within
2x the oracle's own rounding error.
"""
import numpy as np
import pytest

from oracle import frei_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9          # well-conditioned quantities; the contract is 1e-6
RTOL_X = 2e-8        # fluxes vs the 80-bit evaluation of the reference formulas (its own noise ~5e-9)
RTOL_EXACT = 1e-10   # fluxes vs the 40-digit (mpmath) evaluation, on selected wavelength columns
LD = np.longdouble
TINY = 1e-250        # fluxes below this are denormal-range attenuated starlight


@pytest.fixture(params=['auto', 'v2', 'mixed', 'relay'])
def plan(request):
    """Sweep plan: automatic (32-wide warp-chunks, one wavelength per thread, for these small
    cases), forced 64-wide chunks (two per thread: the complete rounds of every production-size
    launch, partly filled last chunk included) and forced mixed (64-wide chunks followed by 32-wide
    ones: what a production-size launch with a short last round runs)."""
    from frei_b200 import _cabi
    lib = _cabi.load()
    _cabi.check(lib.frei_b200_debug_plan({'auto': 0, 'v2': 2, 'mixed': 3, 'relay': 4}[request.param]))
    yield request.param
    _cabi.check(lib.frei_b200_debug_plan(0))


HATCH = []           # one record per comparison: how much of it needed the second clause below


def assert_flux_parity(gpu, ref64, refx, tag=None):
    """
    gpu ~ refx (80-bit evaluation of the reference formulas) to RTOL_X everywhere; gpu ~ ref64 (the
    reference's own fp64 arithmetic) to 1e-6, or — where the fp64 oracle itself is further than
    that from its 80-bit self — within 2x the oracle's own error.  Returns how many elements
    needed that second clause (`n_hatch`), their worst deviation from the fp64 oracle and the
    levels they sit on, and appends the record to HATCH (dumped to gpurun_out/ at session end).
    """
    gpu = np.asarray(gpu, dtype=LD)
    ref64 = np.asarray(ref64, dtype=LD)
    refx = np.asarray(refx, dtype=LD)
    scale = np.maximum(np.abs(refx), LD(TINY))
    ex = np.abs(gpu - refx) / scale
    assert float(ex.max()) < RTOL_X, f'vs extended precision: {float(ex.max()):.3e}'
    e64 = np.abs(gpu - ref64)
    own = np.abs(ref64 - refx)
    strict = e64 <= 1e-6 * scale
    ok = strict | (e64 <= 2 * own + 1e-9 * scale)
    assert bool(ok.all()), f'vs fp64 oracle: {float((e64 / scale).max()):.3e}'
    hatch = ~strict
    rec = dict(tag=tag, n=int(gpu.size), n_hatch=int(hatch.sum()),
               max_rel_vs_fp64=float((e64 / scale).max()),
               max_rel_hatch=float((e64 / scale)[hatch].max()) if hatch.any() else 0.0,
               max_rel_vs_80bit=float(ex.max()),
               oracle_own_error=float((own / scale).max()),
               hatch_levels=sorted(set(np.nonzero(hatch)[0].tolist())) if gpu.ndim == 2 else None)
    HATCH.append(rec)
    return rec


@pytest.fixture(scope='module', autouse=True)
def _dump_hatch_records():
    yield
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if HATCH and os.path.isdir(out):
        with open(os.path.join(out, 'parity_hatch.json'), 'w') as fh:
            json.dump(HATCH, fh, indent=1)


def exact_columns_check(w, tabs_cols, cols, sweeps, gpu_states):
    """
    40-digit evaluation of the reference formulas on wavelength columns `cols`
    through the sweeps [(direction, T_in), ...]; gpu_states[k] = (Fu, Fd) after sweep k.
    """
    import mpmath as mp
    from oracle import frei_oracle_mp as M
    pl = w['planet']
    L = w['L']
    lam_cm = w['lam_um'][cols] * 1e-4
    F_toa = O.F_TOA(lam_cm, pl['T_star'], a_rstar=pl['a_rstar'])
    P = w['P_bar'] * O.BAR
    worst = 0.0
    Fu = [[mp.mpf(0)] * L for _ in cols]
    Fd = [[mp.mpf(0)] * L for _ in cols]
    for (direction, T_in), (gFu, gFd) in zip(sweeps, gpu_states):
        dpg = np.empty(L)
        for i in range(L):
            p2 = P[i] * P[-2] / P[-3] if i == L - 1 else P[i + 1]
            dpg[i] = (P[i] - p2) / pl['g']
        kk = np.stack([O.kappa(tabs_cols, T_in[i], w['P_bar'][i], w['lam_um'][cols], w['mmr'][i],
                               pl['m_bar'])[0] for i in range(L)])
        sig = O.rayleigh_sigma(w['lam_um'][cols], pl['m_bar'])
        for c in range(len(cols)):
            M.sweep_column(direction, kk[:, c], sig[c], dpg, T_in, lam_cm[c], F_toa[c], Fu[c], Fd[c])
            for g, x in ((gFu[:, cols[c]], Fu[c]), (gFd[:, cols[c]], Fd[c])):
                for i in range(L):
                    if abs(x[i]) > TINY:
                        worst = max(worst, float(abs((mp.mpf(float(g[i])) - x[i]) / x[i])))
    return worst


def _rel(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), floor)
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.abs(a - b) / scale
    r[(a == b)] = 0.0
    return r


def _engine(w, dtype=None, T=None, want_dtaus=True, B=1, **kw):
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64
    dtype = dtype or FREI_F64
    tab = synthetic.device_table(w, dtype)
    pl = w['planet']
    T = w['T_init'] if T is None else T
    return Engine(tab, w['lam_um'], np.broadcast_to(w['P_bar'], (B, w['L'])),
                  np.broadcast_to(T, (B, w['L'])), w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                  alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                  want_dtaus=want_dtaus, **kw)


def test_spectral_constants():
    from frei_b200 import synthetic
    w = synthetic.make_workload(10, 3000, 1)
    eng = _engine(w)
    lam_cm = w['lam_um'] * 1e-4
    pl = w['planet']
    # (n^2 - 1) with n - 1 ~ 3e-5 amplifies rounding to ~1e-12 in either implementation
    np.testing.assert_allclose(eng.sigma.cpu().numpy(), O.rayleigh_sigma(w['lam_um'], pl['m_bar']),
                               rtol=1e-11)
    np.testing.assert_allclose(eng.f_toa.cpu().numpy(),
                               O.F_TOA(lam_cm, pl['T_star'], a_rstar=pl['a_rstar']), rtol=1e-13)
    np.testing.assert_allclose(eng.w.cpu().numpy(), O.trapz_weights(lam_cm), rtol=1e-13)
    c1, c2 = eng.c1.cpu().numpy(), eng.c2.cpu().numpy()
    np.testing.assert_allclose(c1 / np.expm1(c2 / 1500.0), O.BB(1500.0, lam_cm), rtol=1e-13)


def test_bracket_bit_exact_vs_scipy():
    """Indices and the out-of-bounds mask must equal scipy's find_indices exactly."""
    from scipy.interpolate._rgi_cython import find_indices
    from frei_b200 import synthetic
    w = synthetic.make_workload(64, 256, 2)
    rs = np.random.RandomState(7)
    axP, axT = w['axis_P'], w['axis_T']
    # on-node, just-off-node, below, above, last node, random
    P = np.concatenate([axP[[0, 3, -1]], np.nextafter(axP[[3, 5]], 0), np.nextafter(axP[[3, 5]], 1e9),
                        [1e-9, 5e4], 10 ** rs.uniform(-7.5, 4.5, 64 - 9)])
    T = np.concatenate([axT[[0, 7, -1]], np.nextafter(axT[[7, 9]], 0), np.nextafter(axT[[7, 9]], 1e9),
                        [100.0, 9600.0], rs.uniform(100, 9600, 64 - 9)])
    w['P_bar'] = P
    eng = _engine(w, T=T)
    dbg = eng.layer_prep(debug=True)
    iP_ref, wP_ref = find_indices((axP,), P[None, :])
    iT_ref, wT_ref = find_indices((axT,), T[None, :])
    oob_ref = (P < axP[0]) | (P > axP[-1]) | (T < axT[0]) | (T > axT[-1])
    for s in range(2):
        assert np.array_equal(dbg['iP'][0, :, s].cpu().numpy(), iP_ref[0])
        assert np.array_equal(dbg['iT'][0, :, s].cpu().numpy(), iT_ref[0])
        assert np.array_equal(dbg['oob'][0, :, s].cpu().numpy().astype(bool), oob_ref)
        assert np.array_equal(dbg['wP'][0, :, s].cpu().numpy(), wP_ref[0])
        assert np.array_equal(dbg['wT'][0, :, s].cpu().numpy(), wT_ref[0])
        for i in range(len(P)):
            ip, wp, op = O.bracket(axP, P[i])
            it, wt, ot = O.bracket(axT, T[i])
            assert (ip, it) == (iP_ref[0, i], iT_ref[0, i]) and (op or ot) == oob_ref[i]


def test_fp64_building_blocks_accuracy():
    """Branch-free 1/x, sqrt, rsqrt, exp(-x), 1-exp(-x) of the kernels: ~1 ulp on their domains."""
    import torch
    from frei_b200 import _cabi
    lib = _cabi.load()
    rs = np.random.RandomState(5)
    x = np.concatenate([10 ** rs.uniform(-12, 3.1, 200000), 10 ** rs.uniform(-300, 300, 20000),
                        np.linspace(0, 2, 4001)[1:], [0.3465735, 0.3465736, 0.6931471, 0.6931472,
                                                       708.0, 745.0, 746.0, 1000.0, 5000.0]])
    d_x = torch.from_numpy(x).cuda()
    out = torch.empty(5 * x.size, dtype=torch.float64, device='cuda')
    _cabi.check(lib.frei_b200_debug_math(d_x.data_ptr(), out.data_ptr(), x.size,
                                         torch.cuda.current_stream().cuda_stream))
    rcp, sq, rsq, ex, om = out.cpu().numpy().reshape(5, -1)
    xl = x.astype(LD)
    ulp = lambda got, ref: float(np.max(np.abs(got.astype(LD) - ref) / np.abs(ref))) / 2.0 ** -52
    assert ulp(rcp, 1 / xl) < 1.5
    assert ulp(sq, np.sqrt(xl)) < 1.5
    assert ulp(rsq, 1 / np.sqrt(xl)) < 2.0
    ok = x < 700                                   # beyond: denormal / zero results
    assert ulp(ex[ok], np.exp(-xl[ok])) < 2.0
    # 1 - exp(-u) is formed as 1 - T for u > ln2/64: up to ~1/u ulp there (< 3e-14 relative)
    assert ulp(om, -np.expm1(-xl)) < 128.0
    # beyond u = 708 the result is flushed to exp(-708) ~ 3.3e-308 (the oracle's values there are
    # denormal or zero: an absolute difference below 1e-300 times the incoming flux)
    big = x > 700
    assert np.all(ex >= 0.0) and np.all(ex[x > 708] <= 3.4e-308)
    assert np.all(np.abs(ex[big] - np.exp(-x[big])) <= 1e-300)


@pytest.mark.parametrize('S', [1, 3, 5, 8])
def test_kappa_matches_interpn(S):
    from frei_b200 import synthetic
    w = synthetic.make_workload(12, 1500, S)
    # include out-of-table levels (fill 0)
    w['P_bar'][0] = 5e4
    T = w['T_init'].copy()
    T[3] = 150.0
    eng = _engine(w, T=T)
    k, sg = eng.kappa()
    k, sg = k[0].cpu().numpy(), sg[0].cpu().numpy()
    tabs = synthetic.host_tables(w)
    pl = w['planet']
    for i in range(w['L']):
        k_ref, s_ref = O.kappa(tabs, T[i], w['P_bar'][i], w['lam_um'], w['mmr'][i], pl['m_bar'])
        np.testing.assert_allclose(k[i], k_ref, rtol=1e-12)
        np.testing.assert_allclose(sg, s_ref, rtol=1e-11)
    assert np.array_equal(k[0], sg) and np.array_equal(k[3], sg)     # fill_value=0 rows


def test_kappa_fp32_table_exact_when_representable():
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_F32
    w = synthetic.make_workload(12, 1500, 3, table_f32=True)
    eng = _engine(w, dtype=FREI_F32)
    k = eng.kappa()[0][0].cpu().numpy()
    tabs = synthetic.host_tables(w)
    for i in range(w['L']):
        k_ref, _ = O.kappa(tabs, w['T_init'][i], w['P_bar'][i], w['lam_um'], w['mmr'][i],
                           w['planet']['m_bar'])
        np.testing.assert_allclose(k[i], k_ref, rtol=1e-13)


def test_propagate_fluxes_both_E_branches():
    from frei_b200.twostream import propagate_fluxes
    rs = np.random.RandomState(3)
    n = 4096
    lam_um = np.logspace(np.log10(0.5), np.log10(10), n)
    F1 = 10 ** rs.uniform(8, 14, n)
    F2 = 10 ** rs.uniform(8, 14, n)
    dtau = 10 ** rs.uniform(-7, 3, n)
    w0 = np.concatenate([10 ** rs.uniform(-9, -1.1, n // 2), rs.uniform(0.1, 0.95, n - n // 2)])
    a, d = propagate_fluxes(lam_um, F1, F2, 1800.0, 1650.0, dtau, omega_0=w0, g_0=0)
    a_ref, d_ref = O.propagate_fluxes(lam_um * 1e-4, F1, F2, 1800.0, 1650.0, dtau, w0, 0)
    a_x, d_x = O.propagate_fluxes(lam_um * 1e-4, F1, F2, 1800.0, 1650.0, dtau.astype(LD), w0, 0)
    assert (w0 > 0.1).sum() > 100 and (w0 <= 0.1).sum() > 100
    assert_flux_parity(a, a_ref, a_x)
    assert_flux_parity(d, d_ref, d_x)


def _oracle_iteration(w, tabs, n_iter, table_kappa=O.kappa, wd=np.float64):
    pl = w['planet']
    lam_cm = w['lam_um'] * 1e-4
    F_toa = O.F_TOA(lam_cm, pl['T_star'], a_rstar=pl['a_rstar']).astype(wd)
    L, n = w['L'], w['n_lam']
    Fu, Fd = np.zeros((L, n), dtype=wd), np.zeros((L, n), dtype=wd)
    T = w['T_init'].copy()
    mmr_fn = (lambda T_, P_, m=w['mmr']: m[0])
    out = []
    for _ in range(n_iter):
        for fn in (O.emit, O.absorb):
            Fu, Fd, T, _, dtaus, dT, bol = fn(tabs, T, w['P_bar'], w['lam_um'], F_toa, pl['g'],
                                              pl['m_bar'], mmr_fn, alpha=pl['alpha'],
                                              fluxes_up=Fu, fluxes_down=Fd, kappa_fn=table_kappa,
                                              work_dtype=wd)
            out.append(dict(Fu=Fu.copy(), Fd=Fd.copy(), T=T.copy(), dtaus=dtaus, dT=dT.copy(),
                            bol=bol))
    return out


@pytest.mark.parametrize('L,n_lam,S,f32', [(20, 1000, 1, False), (50, 5000, 3, False),
                                           (30, 777, 8, False), (24, 1333, 3, True),
                                           (16, 130, 5, False)])
def test_sweeps_match_oracle(L, n_lam, S, f32, plan):
    """Two full emit+absorb iterations: fluxes, dtaus, integrals, dT and T after every sweep."""
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB, FREI_F32, FREI_F64
    w = synthetic.make_workload(L, n_lam, S, table_f32=f32)
    tabs = synthetic.host_tables(w)
    ref = _oracle_iteration(w, tabs, 2)
    refx = _oracle_iteration(w, tabs, 2, wd=LD)
    eng = _engine(w, dtype=FREI_F32 if f32 else FREI_F64)
    k = 0
    gpu_states, worst_x = [], np.zeros(n_lam)
    for it in range(2):
        for direction in (FREI_EMIT, FREI_ABSORB):
            eng.sweep(direction, with_dtaus=True)
            r, rx = ref[k], refx[k]
            k += 1
            Fu, Fd = eng.F_up[0].cpu().numpy(), eng.F_down[0].cpu().numpy()
            gpu_states.append((Fu, Fd))
            tag = f'sweeps L={L} n={n_lam} S={S} f32tab={f32} plan={plan} sweep={k - 1}'
            h_up = assert_flux_parity(Fu, r['Fu'], rx['Fu'], tag + ' F_up')
            h_dn = assert_flux_parity(Fd, r['Fd'], rx['Fd'], tag + ' F_down')
            if k > 1:
                # Only the first emit sweep (zero initial fluxes, optically thin top layers with
                # delta_tau < 1e-6, where the reference's B'/(2E)(chi - psi - xi) grouping cancels)
                # needs the second clause; from the first absorb sweep on, the kernel is within the
                # 1e-6 contract of the reference's own fp64 numbers on every element.
                assert h_up['n_hatch'] == 0 and h_dn['n_hatch'] == 0, (h_up, h_dn)
            else:
                for h in (h_up, h_dn):                  # and there only in the upper atmosphere
                    assert all(lv >= L // 2 for lv in h['hatch_levels']), h
            for g, x in ((Fu, rx['Fu']), (Fd, rx['Fd'])):
                e = np.abs(g.astype(LD) - x) / np.maximum(np.abs(x), LD(TINY))
                worst_x = np.maximum(worst_x, e.max(axis=0).astype(np.float64))
            assert _rel(eng.dtaus[0].cpu().numpy(), r['dtaus']).max() < 1e-12
            sums = eng.sums[0].cpu().numpy()
            lo, hi = (1, L) if direction == FREI_EMIT else (0, L - 1)
            assert _rel(sums[lo:hi], rx['bol'][lo:hi]).max() < 1e-10
            assert _rel(sums[lo:hi], r['bol'][lo:hi]).max() < 1e-7      # fp64 oracle's own noise
            np.testing.assert_allclose(eng.dT[0].cpu().numpy(), r['dT'], rtol=1e-7, atol=1e-8)
            np.testing.assert_allclose(eng.T[0].cpu().numpy(), r['T'], rtol=0, atol=1e-6)
    # arbitration by the exact (40-digit) value of the reference formulas: the columns where the
    # GPU is furthest from the 80-bit oracle plus a few random ones, first emit + absorb
    rs = np.random.RandomState(1)
    cols = np.unique(np.concatenate([np.argsort(worst_x)[-6:], rs.choice(n_lam, 6, replace=False)]))
    tabs_cols = synthetic.host_tables(w, lam_index=cols)
    sweeps = [('emit', w['T_init']), ('absorb', ref[0]['T'])]
    worst = exact_columns_check(w, tabs_cols, cols, sweeps, gpu_states[:2])
    assert worst < RTOL_EXACT, f'vs 40-digit evaluation: {worst:.3e}'


def _one_iteration_with_arbitration(w, align_T):
    """
    emit + absorb on the GPU against the fp64 oracle, the 80-bit oracle and — on the columns where
    GPU and 80-bit oracle disagree most plus a regular sample — the 40-digit evaluation, which
    arbitrates where the reference's grouping is too ill-conditioned even for 80 bits
    (delta_tau down to 1e-12 with omega_0 of order 1).  align_T: start the absorb sweep of the
    GPU from the oracle's temperatures (its own flux noise enters its temperature update).
    """
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    L, n_lam = w['L'], w['n_lam']
    tabs = synthetic.host_tables(w)
    ref = _oracle_iteration(w, tabs, 1)
    refx = _oracle_iteration(w, tabs, 1, wd=LD)
    eng = _engine(w)
    gpu_states, worst_x = [], np.zeros(n_lam)
    for k, direction in enumerate((FREI_EMIT, FREI_ABSORB)):
        if k == 1 and align_T:
            # the fp64 oracle's own flux noise (up to 2e-3 here) enters its temperature update:
            # start the second sweep of both sides from the same temperatures
            eng.set_T(ref[0]['T'])
        eng.sweep(direction, with_dtaus=True)
        Fu, Fd = eng.F_up[0].cpu().numpy(), eng.F_down[0].cpu().numpy()
        gpu_states.append((Fu, Fd))
        for g, x in ((Fu, refx[k]['Fu']), (Fd, refx[k]['Fd'])):
            e = np.abs(g.astype(LD) - x) / np.maximum(np.abs(x), LD(TINY))
            worst_x = np.maximum(worst_x, e.max(axis=0).astype(np.float64))
        assert _rel(eng.dtaus[0].cpu().numpy(), ref[k]['dtaus']).max() < 1e-12
        # the wavelength integrals cannot differ more than the fluxes they sum (rows of mixed sign:
        # relative to the largest integral of the row)
        sums = eng.sums[0].cpu().numpy()
        lo, hi = (1, L) if direction == FREI_EMIT else (0, L - 1)
        bol = np.asarray(refx[k]['bol'], dtype=np.float64)
        floor = 1e-6 * np.abs(bol).max(axis=1, keepdims=True) + 1e-300
        assert _rel(sums[lo:hi], bol[lo:hi], floor=floor[lo:hi]).max() < 1e-10 + 2 * worst_x.max()
        if not align_T:
            np.testing.assert_allclose(eng.T[0].cpu().numpy(), ref[k]['T'], rtol=1e-9, atol=1e-6)
    # Levels outside the table have k = sigma (omega_0 = 1/2) and delta_tau down to 1e-7, where even
    # the 80-bit evaluation of the reference's grouping is noisy; the 40-digit value arbitrates on
    # the columns where GPU and 80-bit oracle disagree most (plus a few others).
    cols = np.unique(np.concatenate([np.argsort(worst_x)[-8:], np.arange(0, n_lam, max(1, n_lam // 6))]))
    tabs_cols = synthetic.host_tables(w, lam_index=cols)
    sweeps = [('emit', w['T_init']), ('absorb', ref[0]['T'])]
    worst = exact_columns_check(w, tabs_cols, cols, sweeps, gpu_states)
    ld_states = [(np.asarray(r['Fu'], dtype=np.float64), np.asarray(r['Fd'], dtype=np.float64)) for r in refx]
    worst_ld = exact_columns_check(w, tabs_cols, cols, sweeps, ld_states)
    assert worst < RTOL_EXACT, f'GPU vs 40-digit evaluation: {worst:.3e} (80-bit oracle: {worst_ld:.3e})'
    assert worst_x.max() < max(RTOL_X, 3 * worst_ld), \
        f'GPU vs 80-bit oracle {worst_x.max():.3e}; 80-bit oracle vs exact {worst_ld:.3e}'
    for k in range(2):                       # the 1e-6 contract against the reference's own fp64 arithmetic
        for g, r64, rx in ((gpu_states[k][0], ref[k]['Fu'], refx[k]['Fu']),
                           (gpu_states[k][1], ref[k]['Fd'], refx[k]['Fd'])):
            scale = np.maximum(np.abs(np.asarray(rx, dtype=LD)), LD(TINY))
            e64 = np.abs(g.astype(LD) - np.asarray(r64, dtype=LD))
            own = np.abs(np.asarray(r64, dtype=LD) - np.asarray(rx, dtype=LD))
            assert bool(((e64 <= 1e-6 * scale) | (e64 <= 2 * own + 1e-8 * scale + 3 * worst_ld * scale)).all())


@pytest.mark.parametrize('L,n_lam,S,T_ref', [(3, 1, 1, 2400.0), (3, 2, 3, 2400.0), (4, 63, 2, 2400.0),
                                             (7, 65, 3, 2400.0), (5, 257, 8, 2400.0),
                                             (12, 510, 3, 9000.0), (12, 129, 2, 120.0)])
def test_ragged_sizes_and_out_of_table_levels(L, n_lam, S, T_ref, plan):
    """
    Smallest legal atmosphere (3 levels), wavelength counts around the warp-chunk and CTA sizes
    (1, 2, 63, 65, 257: odd counts take the one-wavelength-per-thread kernel, the rest the
    two-per-thread one with a partly filled last chunk), and temperature profiles that leave the
    opacity table at the hot / cold end, where the interpolation returns 0 (fill_value=0,
    frei/opacity.py:241-244) for some levels and k falls back to the Rayleigh term alone.
    """
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    w = synthetic.make_workload(L, n_lam, S, T_ref)
    if T_ref != 2400.0:
        inside = (w['T_init'] >= w['axis_T'][0]) & (w['T_init'] <= w['axis_T'][-1])
        assert 0 < inside.sum() < L            # some levels in the table, some outside
    _one_iteration_with_arbitration(w, align_T=(T_ref != 2400.0))


def test_scattering_dominated_mixed_warps(plan):
    """
    Mixing ratios scaled by 1e-3: Rayleigh scattering competes with absorption, omega_0 crosses
    0.1 inside the wavelength range, so warps hold lanes on both branches of E(omega_0)
    (frei/twostream.py:89-94) as well as all-low and all-high warps — the three outcomes of the
    kernel's warp vote.  One emit + absorb iteration against the fp64 / 80-bit oracles.
    """
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    L, n_lam, S = 10, 640, 3
    w = synthetic.make_workload(L, n_lam, S)
    w['mmr'] = w['mmr'] * 1e-3
    tabs = synthetic.host_tables(w)
    pl = w['planet']
    n_mixed = n_low = n_high = 0
    for i in range(L):
        k, sig = O.kappa(tabs, w['T_init'][i], w['P_bar'][i], w['lam_um'], w['mmr'][i], pl['m_bar'])
        hi = ((sig / (sig + k)) > 0.1).reshape(-1, 64)
        n_mixed += int((hi.any(axis=1) & ~hi.all(axis=1)).sum())
        n_low += int((~hi.any(axis=1)).sum())
        n_high += int(hi.all(axis=1).sum())
    assert n_mixed > 5 and n_low > 5 and n_high > 5
    _one_iteration_with_arbitration(w, align_T=True)

def test_reference_kat_and_convergence():
    """
    The reference's own known-answer test (frei/tests/test_core.py:19-71) through
    the mirrored API, plus the converged profile against the oracle (0.1 K).
    """
    import frei_b200 as frei
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet=planet, T_ref=2400)
    op = grid.load_opacities(opacities=frei.load_example_opacity(grid, scale_factor=1))
    assert "1H2-16O" in op
    for attr in ['wavelength', 'temperature', 'pressure']:
        assert hasattr(op.get('1H2-16O'), attr)
    k, sigma = frei.kappa(op, grid.init_temperatures[0], grid.pressures[0], grid.lam,
                          m_bar=planet.m_bar)
    assert np.all(k > sigma)
    assert sigma[0] > sigma[-1]

    # oracle on the same inputs (mock chemistry, as in this container)
    pl = O.hot_jupiter()
    P = O.pressure_grid(30, np.log10(1e-6), np.log10(200))
    T = O.temperature_grid(P, 2400.0, 0.1, 0.1)
    lam, _, _ = O.wavelength_grid(0.5, 10, 500)
    tabs = O.load_example_opacity(P, T, lam, scale_factor=1)
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    k_ref, s_ref = O.kappa(tabs, T[0], P[0], lam, mmr, pl['m_bar'])
    np.testing.assert_allclose(k, k_ref, rtol=1e-12)
    np.testing.assert_allclose(sigma, s_ref, rtol=1e-11)

    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=1)
    s_ref, T_ref, h_ref, d_ref, _ = O.emission_spectrum(tabs, T, P, lam, pl, lambda a, b: mmr,
                                                       n_timesteps=1)
    assert _rel(spec.flux, s_ref).max() < 1e-8
    np.testing.assert_allclose(temps, T_ref, atol=1e-6, rtol=0)
    np.testing.assert_allclose(hist, h_ref, atol=1e-6, rtol=0)
    assert dtaus.shape == d_ref.shape and _rel(dtaus, d_ref).max() < 1e-12
    teff = frei.effective_temperature(grid, spec, dtaus, temps)
    assert abs(teff - O.effective_temperature(P, lam, s_ref, d_ref, T_ref)) < 1e-3

    # full solve: same iteration count, T within 0.1 K (contract), here 1e-3 K
    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=500)
    s_ref, T_ref, h_ref, d_ref, n_it = O.emission_spectrum(tabs, T, P, lam, pl, lambda a, b: mmr,
                                                          n_timesteps=500)
    assert grid.n_iterations == n_it
    assert hist.shape == h_ref.shape
    assert np.abs(temps - T_ref).max() < 1e-3
    assert _rel(spec.flux, s_ref).max() < 1e-6


def test_reference_kat_values_with_fastchem_like_water():
    """
    frei/tests/test_core.py:52-71 pins peak wavelength 1.1518 um +- 0.02, peak flux
    1.296e13 +- 0.1e13 and T_eff 2400 +- 200 K; those values were produced with
    pyfastchem (H2O VMR ~ 3e-4, frei/tests/test_chemistry.py:46).  Feed that
    abundance as the mixing-ratio input and check the GPU path reproduces them.
    """
    import frei_b200 as frei
    from frei_b200 import units as U
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet=planet, T_ref=2400)
    grid.load_opacities(opacities=frei.load_example_opacity(grid, scale_factor=1))
    eng = grid.make_engine(want_dtaus=True)
    eng.set_mmr(3e-4 * 18.0 * U.amu / (2.4 * U.m_p))
    eng.sweep(FREI_EMIT)
    eng.sweep(FREI_ABSORB)
    eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True)
    flux = eng.F_up[0, -1].cpu().numpy()
    lam = np.asarray(grid.lam)
    assert abs(lam[flux.argmax()] - 1.1518) < 0.02
    assert abs(flux.max() - 1.296e13) < 0.1e13
    spec = frei.Spectrum(flux, grid.lam)
    teff = frei.effective_temperature(grid, spec, eng.dtaus[0].cpu().numpy(), eng.T[0].cpu().numpy())
    assert abs(teff - 2400) < 200


def test_emit_absorb_api_host_buffers():
    """emit()/absorb() drop-ins with host arrays: 6-tuple, in-place mutation, dtaus order."""
    import frei_b200 as frei
    from frei_b200 import synthetic
    from frei_b200.opacity import OpacityTable
    w = synthetic.make_workload(12, 400, 1)
    tabs = synthetic.host_tables(w)
    t = tabs['1H2-16O']
    op = {'1H2-16O': OpacityTable(t['values'], t['P'], t['T'], w['lam_um'])}
    pl = w['planet']
    lam_cm = w['lam_um'] * 1e-4
    F_toa = O.F_TOA(lam_cm, pl['T_star'], a_rstar=pl['a_rstar'])
    mmr = O.mock_mmr(['1H2-16O'], pl['m_bar'])
    Fu, Fd = np.zeros((12, 400)), np.zeros((12, 400))
    Fu_r, Fd_r = Fu.copy(), Fd.copy()
    Fu_x, Fd_x = Fu.astype(LD), Fd.astype(LD)
    T = w['T_init']
    for fn, fn_ref in ((frei.emit, O.emit), (frei.absorb, O.absorb)):
        out = fn(op, T, w['P_bar'], w['lam_um'], F_toa, pl['g'] / 100.0, m_bar=pl['m_bar'],
                 n_timesteps=1, alpha=pl['alpha'], fluxes_up=Fu, fluxes_down=Fd)
        ref = fn_ref(tabs, T, w['P_bar'], w['lam_um'], F_toa, pl['g'], pl['m_bar'],
                     lambda a, b: mmr, alpha=pl['alpha'], fluxes_up=Fu_r, fluxes_down=Fd_r)
        fn_ref(tabs, T, w['P_bar'], w['lam_um'], F_toa.astype(LD), pl['g'], pl['m_bar'],
               lambda a, b: mmr, alpha=pl['alpha'], fluxes_up=Fu_x, fluxes_down=Fd_x, work_dtype=LD)
        assert out[0] is Fu and out[1] is Fd                      # mutated in place
        assert_flux_parity(Fu, Fu_r, Fu_x)
        assert_flux_parity(Fd, Fd_r, Fd_x)
        np.testing.assert_allclose(out[2], ref[2], atol=1e-6)
        assert out[3].shape == (12, 2)
        assert _rel(out[4], ref[4]).max() < 1e-12                 # dtaus, visiting order
        np.testing.assert_allclose(out[5], ref[5], rtol=1e-7, atol=1e-9)
        T = out[2]
    # defaults: fluxes None -> F_down[-1] = F_TOA, absorb: F_up[0] = pi B(T0)
    out = frei.absorb(op, w['T_init'], w['P_bar'], w['lam_um'], F_toa, pl['g'] / 100.0,
                      m_bar=pl['m_bar'], n_timesteps=1)
    ref = O.absorb(tabs, w['T_init'], w['P_bar'], w['lam_um'], F_toa, pl['g'], pl['m_bar'],
                   lambda a, b: mmr)
    Fu_x = np.zeros((12, 400), dtype=LD)
    Fu_x[0] = np.pi * O.BB(w['T_init'][0], lam_cm)
    Fd_x = np.zeros((12, 400), dtype=LD)
    Fd_x[-1] = F_toa
    O.absorb(tabs, w['T_init'], w['P_bar'], w['lam_um'], F_toa.astype(LD), pl['g'], pl['m_bar'],
             lambda a, b: mmr, fluxes_up=Fu_x, fluxes_down=Fd_x, work_dtype=LD)
    assert_flux_parity(out[0], ref[0], Fu_x)
    assert_flux_parity(out[1], ref[1], Fd_x)


def test_batch_atmospheres_are_independent(plan):
    """B > 1: every atmosphere of a batch equals its single-atmosphere run, bit for bit."""
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F64
    w = synthetic.make_workload(20, 900, 3)
    tab = synthetic.device_table(w, FREI_F64)
    pl = w['planet']
    B = 5
    rs = np.random.RandomState(0)
    T = w['T_init'][None, :] * rs.uniform(0.7, 1.2, (B, 1))
    g = pl['g'] * rs.uniform(0.5, 2, B)
    fs = rs.uniform(0.5, 2, B)
    mm = w['mmr'][None] * rs.uniform(0.1, 10, (B, 1, 1))

    def run(sel):
        e = Engine(tab, w['lam_um'], np.broadcast_to(w['P_bar'], (len(sel), 20)), T[sel], mm[sel],
                   g=g[sel], m_bar=pl['m_bar'], alpha=1.0, T_star=pl['T_star'],
                   a_rstar=pl['a_rstar'], ftoa_scale=fs[sel])
        for d in (FREI_EMIT, FREI_ABSORB, FREI_EMIT):
            e.sweep(d)
        return e.F_up.cpu().numpy(), e.F_down.cpu().numpy(), e.T.cpu().numpy()
    all_ = run(list(range(B)))
    for b in range(B):
        one = run([b])
        for x, y in zip(all_, one):
            assert np.array_equal(x[b], y[0])


def test_full_size_sampled_parity_and_integrals():
    """
    BASELINE config C2 (50 layers x 200k bins, 3 species) at full size: one
    emit+absorb iteration on the GPU; every wavelength is independent within a
    sweep, so (i) 3000 sampled bins must match the oracle run on just those
    bins, and (ii) the per-layer integrals must equal the trapezoid rule applied
    to the stored fluxes (a checksum of the reduction).
    """
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    L, n_lam, S, T_ref = synthetic.CONFIGS['C2']
    w = synthetic.make_workload(L, n_lam, S, T_ref)
    eng = _engine(w, want_dtaus=False)
    rs = np.random.RandomState(11)
    idx = np.sort(rs.choice(n_lam, 3000, replace=False))
    tabs = synthetic.host_tables(w, lam_index=idx)
    pl = w['planet']
    lam_cm = w['lam_um'] * 1e-4
    F_toa = O.F_TOA(lam_cm[idx], pl['T_star'], a_rstar=pl['a_rstar'])
    Fu_r, Fd_r = np.zeros((L, 3000)), np.zeros((L, 3000))
    Fu_x, Fd_x = Fu_r.astype(LD), Fd_r.astype(LD)
    T = w['T_init'].copy()
    wts = O.trapz_weights(lam_cm)
    for direction, fn in ((FREI_EMIT, O.emit), (FREI_ABSORB, O.absorb)):
        Fu_before = eng.F_up[0].cpu().numpy()
        Fd_before = eng.F_down[0].cpu().numpy()
        eng.sweep(direction)
        Fu, Fd = eng.F_up[0].cpu().numpy(), eng.F_down[0].cpu().numpy()
        fn(tabs, T, w['P_bar'], w['lam_um'][idx], F_toa, pl['g'], pl['m_bar'],
           lambda a, b: w['mmr'][0], alpha=1, fluxes_up=Fu_r, fluxes_down=Fd_r)
        fn(tabs, T, w['P_bar'], w['lam_um'][idx], F_toa.astype(LD), pl['g'], pl['m_bar'],
           lambda a, b: w['mmr'][0], alpha=1, fluxes_up=Fu_x, fluxes_down=Fd_x, work_dtype=LD)
        name = 'emit' if direction == FREI_EMIT else 'absorb'
        h_up = assert_flux_parity(Fu[:, idx], Fu_r, Fu_x, f'C2 full size, 3000 sampled bins, {name} F_up')
        h_dn = assert_flux_parity(Fd[:, idx], Fd_r, Fd_x, f'C2 full size, 3000 sampled bins, {name} F_down')
        if direction == FREI_ABSORB:        # the 1e-6 contract holds element-wise from the first absorb on
            assert h_up['n_hatch'] == 0 and h_dn['n_hatch'] == 0, (h_up, h_dn)
        # integrals: F1_up of step i is fluxes_up[i] (before the sweep for absorb / carried
        # for emit), F1_down = fluxes_down[i] after the sweep.
        sums = eng.sums[0].cpu().numpy()
        lo, hi = (1, L) if direction == FREI_EMIT else (0, L - 1)
        np.testing.assert_allclose(sums[lo:hi, 3], (Fd[lo:hi] * wts).sum(axis=1), rtol=1e-11)
        if direction == FREI_ABSORB:
            np.testing.assert_allclose(sums[lo:hi, 2], (Fu_before[lo:hi] * wts).sum(axis=1), rtol=1e-11)
            np.testing.assert_allclose(sums[lo:hi, 0], (Fu[lo + 1:hi + 1] * wts).sum(axis=1), rtol=1e-11)
        else:
            np.testing.assert_allclose(sums[1:L - 1, 1], (Fd_before[2:L] * wts).sum(axis=1), rtol=1e-11)
        T = eng.T[0].cpu().numpy()      # continue the oracle from the GPU's T (global integrals)


def test_gpu_vs_reference_own_run():
    """
    The GPU path through the mirrored API against tests/golden/reference_run.json — outputs of
    the reference's OWN source files run under dependency stubs (tests/golden/run_reference.py).
    """
    import json
    import os
    import frei_b200 as frei
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_run.json')) as fh:
        ref = json.load(fh)
    a = ref['A']
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet=planet, T_ref=2400)
    op = grid.load_opacities(opacities=frei.load_example_opacity(grid, scale_factor=1))
    idx = a['lam_index']
    k, sigma = frei.kappa(op, grid.init_temperatures[0], grid.pressures[0], grid.lam, m_bar=planet.m_bar)
    np.testing.assert_allclose(np.asarray(k)[idx], a['kappa0'], rtol=1e-12)
    np.testing.assert_allclose(np.asarray(sigma)[idx], a['sigma'], rtol=1e-11)
    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=1)
    # 1e-6 is the contract; the reference's own fp64 noise in thin layers is ~1e-9 here
    np.testing.assert_allclose(np.asarray(spec.flux)[idx], a['spectrum'], rtol=1e-8)
    np.testing.assert_allclose(temps, a['final_temps'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(hist, a['temp_hist'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(dtaus[:, idx], a['dtaus'], rtol=1e-11)
    assert abs(frei.effective_temperature(grid, spec, dtaus, temps) - a['T_eff']) < 1e-4
    # full solve on the small grid: same number of iterations, converged T within 0.1 K (contract)
    b = ref['B']
    grid = frei.Grid(planet=planet, T_ref=2400, n_layers=12, n_wl_bins=120)
    grid.load_opacities(opacities=frei.load_example_opacity(grid, scale_factor=1))
    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=400)
    assert hist.shape[1] == b['n_columns']
    assert np.abs(np.asarray(temps) - np.array(b['final_temps'])).max() < 1e-3
    np.testing.assert_allclose(np.asarray(spec.flux), b['spectrum'], rtol=1e-6)
    # propagate_fluxes, both E branches
    c = ref['C']
    F2u, F1d = frei.propagate_fluxes(np.array(c['lam_um']), np.array(c['F1']), np.array(c['F2']),
                                     1800.0, 1650.0, np.array(c['dtau']), omega_0=np.array(c['w0']), g_0=0)
    np.testing.assert_allclose(F2u, c['F2u'], rtol=1e-9)
    np.testing.assert_allclose(F1d, c['F1d'], rtol=1e-9)


def test_batch_solve_device_convergence_matches_per_atmosphere_oracle():
    """
    A batch of different atmospheres (T profile, gravity, irradiation, abundances) solved in
    lock-step with the convergence rule evaluated on the device: every atmosphere must stop after
    the same number of iterations as the oracle's Grid.emission_spectrum and end at the same T.
    """
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64
    L, n_lam, S = 14, 600, 3
    w = synthetic.make_workload(L, n_lam, S)
    tab = synthetic.device_table(w, FREI_F64)
    tabs = synthetic.host_tables(w)
    pl = w['planet']
    B = 5
    rs = np.random.RandomState(4)
    T0 = w['T_init'][None, :] * np.array([0.6, 0.8, 1.0, 1.1, 0.9])[:, None]
    g = pl['g'] * np.array([0.5, 1.0, 2.0, 4.0, 1.5])
    a_rstar = pl['a_rstar'] * np.array([0.8, 1.0, 1.5, 2.5, 1.2])
    mm = w['mmr'][None] * np.array([0.1, 1.0, 10.0, 3.0, 0.3])[:, None, None]
    eng = Engine(tab, w['lam_um'], np.broadcast_to(w['P_bar'], (B, L)), T0, mm, g=g,
                 m_bar=pl['m_bar'], alpha=1.0, T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                 ftoa_scale=(pl['a_rstar'] / a_rstar) ** 2)
    iters, T = eng.solve_batch(300, check_every=3)
    spec = eng.F_up[:, L - 1, :].cpu().numpy()
    assert len(set(iters.tolist())) > 1                        # they really stop at different times
    for b in range(B):
        planet = dict(pl, g=g[b], a_rstar=a_rstar[b])
        s_ref, T_ref, hist, dtaus, n_it = O.emission_spectrum(
            tabs, T0[b], w['P_bar'], w['lam_um'], planet, lambda x, y, m=mm[b, 0]: m, n_timesteps=300)
        assert iters[b] == n_it, (b, iters[b], n_it)
        assert np.abs(T[b] - T_ref).max() < 1e-3
        assert _rel(spec[b], s_ref).max() < 1e-6


def test_groupby_bins_agg_matches_reference_binning():
    """Wavelength binning kernel vs the restated numba loop (frei/interp.py:174-202)."""
    from frei_b200.interp import groupby_bins_agg, cut_codes
    from oracle import binning_oracle as BO
    import pandas as pd
    rs = np.random.RandomState(8)
    lam_edges = np.concatenate([[0.45], np.logspace(np.log10(0.5), 1, 300)])
    # cut codes == pandas.cut codes, including values on the edges and outside
    probe = np.concatenate([lam_edges, np.nextafter(lam_edges, 0), np.nextafter(lam_edges, 100),
                            rs.uniform(0.3, 12, 500)])
    assert np.array_equal(cut_codes(probe, lam_edges), np.asarray(pd.cut(probe, lam_edges).codes))
    for n, dt in ((20000, np.float64), (150000, np.float32), (3100, np.float64)):
        wl = np.sort(rs.uniform(0.46, 9.99, n))               # inside the outer edges, like the crop
        a = (10 ** rs.uniform(-3, 2, (3, 4, n))).astype(dt)
        out = groupby_bins_agg(a, wl, lam_edges, func=np.trapezoid)
        ref, centres = BO.groupby_bins_agg(a, wl, lam_edges)
        assert out.shape == (3, 4, 300)
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(out.wavelength, centres, rtol=1e-15)
    # unsorted samples: several runs per bin, and isolated out-of-range samples are skipped
    wl = rs.uniform(0.46, 9.99, 5000)
    wl[::97] = 20.0
    a = rs.uniform(0, 1, (2, 5000))
    out = groupby_bins_agg(a, wl, lam_edges)
    ref, _ = BO.groupby_bins_agg(a, wl, lam_edges)
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-300)
    wl[10:12] = 20.0                                          # two consecutive outside -> error
    with pytest.raises(ValueError, match='negative indices'):
        groupby_bins_agg(a, wl, lam_edges)
    with pytest.raises(ValueError, match='negative indices'):
        BO.groupby_bins_agg(a, wl, lam_edges)


def test_binned_opacity_from_bin_directory(tmp_path):
    """HELIOS-K .bin directory -> binned, regridded tables (frei/opacity.py:66-170, 395-483)."""
    import frei_b200 as frei
    from frei_b200.opacity import read_opacity_dir, binned_opacity
    from oracle import binning_oracle as BO
    rs = np.random.RandomState(3)
    d = tmp_path / '1H2-16O__POKAZATEL_e2b'
    d.mkdir()
    w0, w1 = 900, 21000                                  # cm^-1  ->  0.476 .. 11.1 micron
    n = int(round((w1 - w0) / 0.01))
    raw = {}
    for T in (500, 1500, 2500):
        for ptag, P in (('n300', 1e-3), ('p000', 1.0), ('p200', 100.0)):
            x = (10 ** rs.uniform(-6, 1, n)).astype(np.float32)
            x.tofile(str(d / f'Out_{w0:05d}_{w1:05d}_{T:05d}_{ptag}.bin'))
            raw[(T, P)] = x
    T_ax, P_ax, wl, grid = read_opacity_dir(str(d))
    assert list(T_ax) == [500, 1500, 2500] and np.allclose(P_ax, [1e-3, 1.0, 100.0])
    assert grid.shape == (3, 3, n - 1) and np.all(np.diff(wl) > 0)
    assert np.array_equal(grid[1, 2], raw[(1500, 100.0)][1:][::-1])
    planet = frei.Planet.from_hot_jupiter()
    g = frei.Grid(planet, n_wl_bins=150, n_layers=9, T_ref=1800)
    tabs = binned_opacity(g.init_temperatures, g.pressures, g.wl_bins, g.lam, species=['H2O'],
                          path=str(tmp_path / '*'))
    tab = tabs['1H2-16O']
    ref, centres = BO.binned_opacity_one(grid, wl, T_ax, P_ax, np.asarray(g.init_temperatures),
                                         np.asarray(g.pressures), np.asarray(g.wl_bins))
    got = np.transpose(np.asarray(tab.values), (1, 0, 2))         # -> [T, P, wavelength]
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(tab.wavelength, centres, rtol=1e-14)
    # the loaded tables feed the hot path
    g.load_opacities(opacities=tabs)
    spec, temps, hist, dtaus = g.emission_spectrum(n_timesteps=2)
    assert np.all(np.isfinite(np.asarray(spec.flux))) and dtaus.shape == (9, 150)


@pytest.mark.parametrize('L,n_lam,S', [(50, 5000, 3), (30, 1026, 8), (24, 1333, 2)])
def test_fp32_mode_sweeps_within_1e4(L, n_lam, S, plan):
    """
    fp32 arithmetic mode (flux state + table in fp32, integrals in fp64): per-wavelength fluxes
    within 1e-4 relative of the fp64 oracle (BASELINE.json north_star) through three RE iterations.
    Values more than 30 decades below the row maximum are outside fp32's range and compared
    absolutely.
    """
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB, FREI_F32
    w = synthetic.make_workload(L, n_lam, S, table_f32=True)
    tabs = synthetic.host_tables(w)
    ref = _oracle_iteration(w, tabs, 3)
    eng = _engine(w, dtype=FREI_F32, flux_dtype=FREI_F32)
    k = 0
    worst = 0.0
    for it in range(3):
        for direction in (FREI_EMIT, FREI_ABSORB):
            eng.sweep(direction, with_dtaus=True)
            r = ref[k]
            k += 1
            for g, x in ((eng.F_up[0], r['Fu']), (eng.F_down[0], r['Fd'])):
                g = g.cpu().numpy().astype(np.float64)
                scale = np.maximum(np.abs(x), 1e-30 * np.abs(x).max(axis=1, keepdims=True) + 1e-300)
                worst = max(worst, float((np.abs(g - x) / scale).max()))
            assert _rel(eng.dtaus[0].cpu().numpy().astype(np.float64), r['dtaus']).max() < 1e-5
            lo, hi = (1, L) if direction == FREI_EMIT else (0, L - 1)
            assert _rel(eng.sums[0].cpu().numpy()[lo:hi], r['bol'][lo:hi]).max() < 1e-5
            np.testing.assert_allclose(eng.T[0].cpu().numpy(), r['T'], rtol=0, atol=0.05)
    assert worst < 1e-4, f'fp32 flux error {worst:.3e}'


# ---------------------------------------------------------------------------
# device-side post-processing: T_eff sums, Milne pressure, contribution function
# ---------------------------------------------------------------------------
@pytest.mark.parametrize('L,n_lam', [(2, 7), (3, 33), (4, 300), (5, 257), (9, 1000), (30, 5000), (100, 777)])
def test_diagnostics_kernel_matches_numpy_interp_on_unsorted_columns(L, n_lam):
    """
    frei/core.py:392-395 calls np.interp on exp(-dtaus[:, j]), which is not sorted in general
    (row 0 is the row of ones); the kernel restates numpy's search path.  Random unsorted
    columns exercise the linear (len <= 4), guess and bisection branches.
    """
    import torch
    from frei_b200.core import _device_diagnostics, _trapz_weights_cm
    rng = np.random.default_rng(100 + L)
    dtaus = 10.0 ** rng.uniform(-6, 1.5, (L, n_lam))
    dtaus[0] = 1.0
    kind = rng.integers(0, 3, n_lam)
    mono = np.sort(dtaus[1:], axis=0)[::-1]                 # decreasing upwards, like a real solve
    dtaus[1:, kind == 1] = mono[:, kind == 1]
    dtaus[1:, kind == 2] = mono[:, kind == 2] * rng.uniform(0.5, 2.0, (L - 1, int((kind == 2).sum())))
    lam = np.logspace(np.log10(0.5), 1, n_lam)
    P = np.logspace(-6, np.log10(200), L)[::-1].copy()
    T = 2400.0 * (P / 0.1) ** 0.1
    spec = 10.0 ** rng.uniform(8, 13, n_lam)
    w = _trapz_weights_cm(lam)
    sums, pm, cf = _device_diagnostics(lam, w, P, T, spec, dtaus, want_pressure=True, want_cf=True)
    torch.cuda.synchronize()
    pm_ref = O.pressure_milne(P, dtaus)
    np.testing.assert_allclose(pm.cpu().numpy(), pm_ref, rtol=1e-12, atol=0)
    sums = sums.cpu().numpy()
    wt = spec * lam * 1e-4
    np.testing.assert_allclose(sums, [(pm_ref * wt).sum(), wt.sum(), (w * spec).sum()], rtol=1e-12)
    cf_ref = O.contribution_function(lam, P, T, dtaus)
    cf = cf.cpu().numpy()
    assert np.abs(cf.sum(axis=0) - 1).max() < 1e-12
    np.testing.assert_allclose(cf, cf_ref, rtol=1e-10, atol=1e-300)


def test_grid_diagnostics_from_resident_state():
    """Grid.diagnostics(): T_eff and contribution function of a solve without host copies."""
    import frei_b200 as frei
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet=planet, T_ref=2400)
    grid.load_opacities(opacities=frei.load_example_opacity(grid, scale_factor=1))
    spec, temps, hist, dtaus = grid.emission_spectrum(n_timesteps=3)
    d = grid.diagnostics(contribution_function=True, pressure_milne=True)
    P, lam = np.asarray(grid.pressures), np.asarray(grid.lam)
    flux, temps = np.asarray(spec.flux), np.asarray(temps)
    assert abs(d['T_milne'] - O.effective_temperature_milne(P, lam, flux, dtaus, temps)) < 1e-6
    assert abs(d['T_planck'] - O.effective_temperature_planck(lam, flux)) < 1e-6
    assert abs(d['T_eff'] - O.effective_temperature(P, lam, flux, dtaus, temps)) < 1e-6
    assert abs(d['T_eff'] - frei.effective_temperature(grid, spec, dtaus, temps)) < 1e-9
    np.testing.assert_allclose(d['pressure_milne'], O.pressure_milne(P, dtaus), rtol=1e-12)
    np.testing.assert_allclose(d['contribution_function'], O.contribution_function(lam, P, temps, dtaus),
                               rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(frei.contribution_function(grid, dtaus, temps), d['contribution_function'],
                               rtol=1e-13, atol=1e-300)


def test_split_reduce_update_sequence_equals_fused_post(plan):
    """
    The three-call sequence of the NCCL mode on one GPU — frei_b200_sweep, frei_b200_reduce,
    (all-reduce of ws->sums would go here,) frei_b200_update_T with the records rebuilt in the same
    launch — against the fused frei_b200_sweep_step: same reduction order and the same update
    code, so T, dT, the integrals and the fluxes agree to rounding after two iterations.
    """
    import ctypes as C
    import torch
    from frei_b200 import synthetic, _cabi
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    w = synthetic.make_workload(20, 1000, 3)
    fused, split = _engine(w, want_dtaus=False), _engine(w, want_dtaus=False)
    lib = split.lib
    st = split._stream()
    split.layer_prep()
    flux = split._flux_struct(False)
    for it in range(2):
        for direction in (FREI_EMIT, FREI_ABSORB):
            fused.sweep(direction)
            _cabi.check(lib.frei_b200_sweep(C.byref(split._tab), C.byref(split._spec), C.byref(split._atm),
                                            C.byref(flux), direction, C.byref(split._ws), st))
            _cabi.check(lib.frei_b200_reduce(C.byref(split._atm), C.byref(split._ws), split.n_lam, st))
            _cabi.check(lib.frei_b200_update_T(C.byref(split._tab), C.byref(split._atm), C.byref(split._ws),
                                               direction, -1.0, None, st))
            torch.cuda.synchronize()
            for name in ('T', 'dT', 'sums', 'F_up', 'F_down'):
                a, b = getattr(fused, name).cpu().numpy(), getattr(split, name).cpu().numpy()
                np.testing.assert_allclose(b, a, rtol=1e-12, atol=1e-300, err_msg=f'{name}, iteration {it}')


@pytest.mark.parametrize('L,n_lam,S,force', [(100, 2050, 8, 2), (100, 2050, 8, 3), (100, 2050, 8, 4),
                                             (200, 3002, 3, 2), (200, 3002, 3, 3), (200, 3002, 3, 4)])
def test_large_shapes_match_oracle(L, n_lam, S, force):
    """
    The layer counts and species counts of BASELINE configs C3 (100 layers, 8 species) and C5
    (200 layers), at a few thousand wavelengths, with 64-wide chunks (force 2) and the mixed plan
    (force 3): one emit + absorb iteration against the fp64 and 80-bit oracle.
    """
    from frei_b200 import synthetic, _cabi
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    lib = _cabi.load()
    _cabi.check(lib.frei_b200_debug_plan(force))
    try:
        w = synthetic.make_workload(L, n_lam, S, 3200.0 if S == 8 else 2400.0)
        tabs = synthetic.host_tables(w)
        ref = _oracle_iteration(w, tabs, 1)
        refx = _oracle_iteration(w, tabs, 1, wd=LD)
        eng = _engine(w)
        for k, direction in enumerate((FREI_EMIT, FREI_ABSORB)):
            eng.sweep(direction, with_dtaus=True)
            Fu, Fd = eng.F_up[0].cpu().numpy(), eng.F_down[0].cpu().numpy()
            tag = f'large L={L} n={n_lam} S={S} plan={force} sweep={k}'
            assert_flux_parity(Fu, ref[k]['Fu'], refx[k]['Fu'], tag + ' F_up')
            assert_flux_parity(Fd, ref[k]['Fd'], refx[k]['Fd'], tag + ' F_down')
            assert _rel(eng.dtaus[0].cpu().numpy(), ref[k]['dtaus']).max() < 1e-12
            sums = eng.sums[0].cpu().numpy()
            lo, hi = (1, L) if direction == FREI_EMIT else (0, L - 1)
            assert _rel(sums[lo:hi], refx[k]['bol'][lo:hi]).max() < 1e-10
            np.testing.assert_allclose(eng.dT[0].cpu().numpy(), ref[k]['dT'], rtol=1e-6, atol=1e-8)
            np.testing.assert_allclose(eng.T[0].cpu().numpy(), ref[k]['T'], rtol=0, atol=1e-6)
    finally:
        _cabi.check(lib.frei_b200_debug_plan(0))


def test_maximum_layer_count():
    """
    256 levels, the most the reduction supports (4 L = 1024 threads; 257 is rejected).  With 256
    levels between 1e-6 and 200 bar the layers are so thin (delta_tau down to 1e-9) that even the
    80-bit evaluation of the reference's grouping is noisy, so the 40-digit evaluation arbitrates.
    """
    from frei_b200 import synthetic, _cabi
    from frei_b200._cabi import FreiError
    lib = _cabi.load()
    _cabi.check(lib.frei_b200_debug_plan(2))
    try:
        _one_iteration_with_arbitration(synthetic.make_workload(256, 514, 1), align_T=True)
        with pytest.raises(FreiError, match='bad argument|levels'):
            from frei_b200.engine import FREI_EMIT
            _engine(synthetic.make_workload(257, 128, 1)).sweep(FREI_EMIT)
    finally:
        _cabi.check(lib.frei_b200_debug_plan(0))


def test_c4_shaped_batch_against_per_atmosphere_oracle():
    """
    A C4-shaped batch (T_eq x log g x metallicity grid, here 8 x 8 x 4 = 256 atmospheres, 20 layers
    x 192 bins) solved with the device-side convergence rule.  A spread sample of atmospheres is
    solved one by one with the oracle's Grid.emission_spectrum: same iteration count, T within
    1e-3 K, spectrum within 1e-6.  The cold, metal-rich corner diverges in the reference's explicit
    scheme (T < 0, then NaN): there the rule must still stop the atmosphere — np.sign(nan) !=
    np.sign(nan) counts as a zero crossing in frei/core.py:306-311 — and both end with NaN.
    """
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64
    L, n_lam, S = 20, 192, 3
    w = synthetic.make_workload(L, n_lam, S)
    tab = synthetic.device_table(w, FREI_F64)
    tabs = synthetic.host_tables(w)
    pl = w['planet']
    tt, gg, mm = [x.ravel() for x in np.meshgrid(np.linspace(1000, 2500, 8), np.linspace(2.5, 4.0, 8),
                                                 np.linspace(-1, 2, 4), indexing='ij')]
    B = tt.size
    T0 = tt[:, None] * (w['P_bar'][None, :] / 0.1) ** 0.1
    mmr = w['mmr'][None] * (10.0 ** mm)[:, None, None]
    eng = Engine(tab, w['lam_um'], np.broadcast_to(w['P_bar'], (B, L)), T0, mmr, g=10.0 ** gg,
                 m_bar=pl['m_bar'], alpha=1.0, T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                 ftoa_scale=(tt / 2400.0) ** 4)
    cap = 300
    iters, T = eng.solve_batch(cap, check_every=5)
    spec = eng.F_up[:, L - 1, :].cpu().numpy()
    assert (iters < cap).all(), f'{(iters >= cap).sum()} atmospheres were stopped by the cap'
    finite = np.isfinite(T).all(axis=1)
    sample = sorted(set(np.linspace(0, B - 1, 14).astype(int).tolist() + np.flatnonzero(~finite)[:3].tolist()))
    n_nan = 0
    for b in sample:
        planet = dict(pl, g=10.0 ** gg[b], a_rstar=pl['a_rstar'] / (tt[b] / 2400.0) ** 2)
        with np.errstate(all='ignore'):
            s_ref, T_ref, hist, dtaus, n_it = O.emission_spectrum(
                tabs, T0[b], w['P_bar'], w['lam_um'], planet, lambda x, y, m=mmr[b, 0]: m, n_timesteps=cap)
        if np.isfinite(T_ref).all():
            assert finite[b] and iters[b] == n_it, (b, iters[b], n_it)
            assert np.abs(T[b] - T_ref).max() < 1e-3
            assert _rel(spec[b], s_ref).max() < 1e-6
        else:
            n_nan += 1
            assert not finite[b], b                        # diverged on both sides ...
            assert n_it < 12 and iters[b] < 12, (b, iters[b], n_it)   # ... and stopped by the rule at once
    assert n_nan >= 1 or finite.all()
