"""
Memory-safety and race evidence without compute-sanitizer (the tool is closed on this GPU pool:
profiles/sanitizer/r02_compute_sanitizer_closed.log).  Two properties a memcheck / racecheck
failure would break are checked directly, on every sweep plan and kernel family:

* guard bands: every buffer a kernel writes (flux state, dtaus, level records, partial rows +
  chunk sums + tickets + plan header, sums, T history, dT) is carved out of an arena filled with a
  sentinel, with guard bands before and after it; after the runs the bands must be untouched
  (out-of-bounds writes) and the outputs must contain no sentinel (reads of memory nobody wrote
  would propagate its NaN payload into the results);
* repeatability: the same sweeps run again from the same state give bit-identical fluxes,
  integrals and temperatures — the hand-rolled pieces (thread-private cp.async slots, the
  mbarrier/TMA prologue, programmatic dependent launch between sweep and post kernel, the ticket
  counter of the two-stage reduction) have no ordering freedom that changes a result.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096                                   # doubles on either side of every carved buffer
SENTINEL = float(np.frombuffer(np.uint64(0x7ff8dead0000beef).tobytes(), dtype=np.float64)[0])   # a NaN payload


def _carve(arena, offset, like):
    """A view of `arena` shaped like `like`, starting GUARD doubles after `offset`; returns (view, next offset)."""
    import torch
    n_bytes = like.numel() * like.element_size()
    n_dbl = ((n_bytes + 255) // 256) * 32          # whole 256-byte lines: buffers stay 256-byte aligned
    start = offset + GUARD
    raw = arena[start:start + n_dbl]
    view = raw.view(torch.uint8)[:n_bytes].view(like.dtype).view(like.shape)
    view.copy_(like)
    return view, start + n_dbl


def _guarded_engine(w, B=1, flux_dtype=None, want_dtaus=True):
    """Engine whose kernel-written buffers live between guard bands of one sentinel-filled arena."""
    import torch
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64, FREI_F32
    fd = FREI_F64 if flux_dtype is None else flux_dtype
    table = synthetic.device_table(w, FREI_F32 if fd == FREI_F32 else FREI_F64)
    pl = w['planet']
    L, S = w['L'], w['S']
    T0 = np.broadcast_to(w['T_init'], (B, L)) * np.linspace(0.85, 1.1, B)[:, None]
    eng = Engine(table, w['lam_um'], np.broadcast_to(w['P_bar'], (B, L)), T0, np.broadcast_to(w['mmr'], (B, L, S)),
                 g=pl['g'], m_bar=pl['m_bar'], alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                 flux_dtype=fd, want_dtaus=want_dtaus)
    names = ['F_up', 'F_down', 'dtaus', '_lp', '_partials', 'sums', 'hist', 'T']
    bufs = [getattr(eng, n) for n in names if getattr(eng, n) is not None]
    total = sum(((b.numel() * b.element_size() + 255) // 256) * 32 + GUARD for b in bufs) + GUARD
    arena = torch.full((total,), SENTINEL, dtype=torch.float64, device=eng.device)
    off, spans = 0, []
    for n in names:
        b = getattr(eng, n)
        if b is None:
            continue
        view, end = _carve(arena, off, b)
        spans.append((off, off + GUARD))              # the band in front of this buffer
        setattr(eng, n, view)
        off = end
    spans.append((off, off + GUARD))
    eng.dT = eng.hist[2]
    eng._records_stale = True
    eng._build_structs()
    return eng, arena, spans


def _bands_intact(arena, spans):
    import torch
    bits = arena.view(torch.int64)
    want = torch.tensor([SENTINEL], dtype=torch.float64, device=arena.device).view(torch.int64)
    return all(bool((bits[a:b] == want).all()) for a, b in spans)


def _has_sentinel(t):
    import torch
    if t.dtype != torch.float64:
        return False
    want = torch.tensor([SENTINEL], dtype=torch.float64, device=t.device).view(torch.int64)
    return bool((t.contiguous().view(torch.int64) == want).any())


@pytest.mark.parametrize('plan', [0, 1, 2, 3, 4])
@pytest.mark.parametrize('L,n_lam,S,B', [(12, 1000, 3, 1), (20, 514, 8, 1), (9, 333, 1, 1), (7, 260, 5, 1),
                                         (10, 600, 3, 3), (50, 8192, 3, 1)])
def test_guard_bands_and_no_uninitialised_reads(plan, L, n_lam, S, B):
    import torch
    from frei_b200 import synthetic, _cabi
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    lib = _cabi.load()
    _cabi.check(lib.frei_b200_debug_plan(plan))
    try:
        w = synthetic.make_workload(L, n_lam, S)
        eng, arena, spans = _guarded_engine(w, B=B)
        for _ in range(2):
            eng.sweep(FREI_EMIT, T_hist=eng.hist[0])
            eng.sweep(FREI_ABSORB, T_hist=eng.hist[1])
        eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True)
        torch.cuda.synchronize()
        assert _bands_intact(arena, spans), 'a kernel wrote outside its buffer'
        for name in ('F_up', 'F_down', 'dtaus', 'sums', 'T', 'hist'):
            assert not _has_sentinel(getattr(eng, name)), f'{name} carries never-written memory'
        assert torch.isfinite(eng.T).all() and torch.isfinite(eng.F_up).all()
    finally:
        _cabi.check(lib.frei_b200_debug_plan(0))


@pytest.mark.parametrize('n_lam', [1024, 333, 514])
def test_guard_bands_fp32_sweep(n_lam):
    import torch
    from frei_b200 import synthetic
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB, FREI_F32
    w = synthetic.make_workload(10, n_lam, 3, table_f32=True)
    eng, arena, spans = _guarded_engine(w, flux_dtype=FREI_F32)
    for _ in range(2):
        eng.sweep(FREI_EMIT)
        eng.sweep(FREI_ABSORB)
    eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True)
    torch.cuda.synchronize()
    assert _bands_intact(arena, spans)
    assert torch.isfinite(eng.T).all() and torch.isfinite(eng.F_up.double()).all()
    assert torch.isfinite(eng.dtaus.double()).all()


@pytest.mark.parametrize('plan', [0, 2, 3, 4])
def test_repeated_runs_are_bit_identical(plan):
    """20 repetitions of two RE iterations from the same state, production-size grid (all SMs busy,
    several rounds of chunks, the reduction's ticket decided by a different CTA every time)."""
    import torch
    from frei_b200 import synthetic, _cabi
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    lib = _cabi.load()
    _cabi.check(lib.frei_b200_debug_plan(plan))
    try:
        w = synthetic.make_workload(30, 160_000, 3)
        eng, arena, spans = _guarded_engine(w, want_dtaus=False)
        first = None
        for rep in range(20):
            eng.reset(w['T_init'])
            for _ in range(2):
                eng.sweep(FREI_EMIT)
                eng.sweep(FREI_ABSORB)
            torch.cuda.synchronize()
            state = (eng.F_up.clone(), eng.F_down.clone(), eng.sums.clone(), eng.T.clone())
            if first is None:
                first = state
            else:
                for a, b in zip(first, state):
                    assert torch.equal(a, b), f'repetition {rep} differs'
        assert _bands_intact(arena, spans)
    finally:
        _cabi.check(lib.frei_b200_debug_plan(0))


@pytest.mark.parametrize('L,n_lam,S', [(30, 160_000, 3), (50, 200_000, 3), (12, 1000, 3), (100, 40_000, 8), (9, 4098, 1)])
def test_relay_plan_is_bit_identical_to_whole_chunks(L, n_lam, S):
    """The relay plan (automatic at 160k and 200k bins: 1.06 and 1.32 chunks per resident warp; forced
    at the small sizes) cuts chunks between warps but computes every (chunk, layer) with the same
    instructions as the plan of whole 64-wide chunks: fluxes, dtaus, integrals and T must not differ
    in a single bit, sweep after sweep."""
    import torch
    from frei_b200 import synthetic, _cabi
    from frei_b200.engine import FREI_EMIT, FREI_ABSORB
    lib = _cabi.load()
    w = synthetic.make_workload(L, n_lam, S)
    states = {}
    try:
        for plan in (2, 4 if n_lam < 100_000 else 0):
            _cabi.check(lib.frei_b200_debug_plan(plan))
            eng, arena, spans = _guarded_engine(w)
            snap = []
            for _ in range(2):
                eng.sweep(FREI_EMIT)
                snap += [eng.F_up.clone(), eng.F_down.clone(), eng.sums.clone(), eng.T.clone()]
                eng.sweep(FREI_ABSORB)
                snap += [eng.F_up.clone(), eng.F_down.clone(), eng.sums.clone(), eng.T.clone()]
            eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True)
            torch.cuda.synchronize()
            snap += [eng.F_up.clone(), eng.F_down.clone(), eng.dtaus.clone(), eng.T.clone()]
            assert _bands_intact(arena, spans)
            states[plan] = snap
            del eng, arena
        (pa, a), (pb, b) = states.items()
        for k, (x, y) in enumerate(zip(a, b)):
            assert torch.equal(x, y), f'plan {pa} vs {pb}: snapshot {k} differs'
    finally:
        _cabi.check(lib.frei_b200_debug_plan(0))


def test_batch_tracker_and_diagnostics_guarded():
    import torch
    from frei_b200 import synthetic
    w = synthetic.make_workload(10, 600, 3)
    eng, arena, spans = _guarded_engine(w, B=5)
    iters, T = eng.solve_batch(8, check_every=2)
    torch.cuda.synchronize()
    assert _bands_intact(arena, spans)
    assert np.isfinite(T).all()


@pytest.mark.parametrize('n_lam', [160_000, 20_000])
def test_cuda_graph_replay_equals_plain_launches(n_lam):
    """An RE iteration captured into a CUDA graph (emit, post, absorb, post) and replayed gives bit for
    bit what the individual launches give — at 160k bins with the relay plan, whose hand-over flags
    must be back at zero after every launch because a replay cannot change a kernel argument — and
    a change of the engine's structs (the batch tracker) drops the captured graph instead of
    replaying stale pointers."""
    import torch
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64
    w = synthetic.make_workload(30, n_lam, 3)
    tab = synthetic.device_table(w, FREI_F64)
    pl = w['planet']

    def mk():
        return Engine(tab, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'], m_bar=pl['m_bar'],
                      alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'])
    a, b = mk(), mk()
    assert b.capture_iteration()
    for _ in range(4):
        a.iteration()
        b.iteration()
    torch.cuda.synchronize()
    assert b._graph is not None
    for name in ('T', 'F_up', 'F_down', 'sums'):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.isfinite(b.T).all()
    # structs change -> the graph is dropped, the tracker is honoured
    b.enable_batch_convergence(n_zero_crossings=10 ** 9, convergence_dT=0.0)
    assert b._graph is None
    a.enable_batch_convergence(n_zero_crossings=10 ** 9, convergence_dT=0.0)
    for _ in range(2):
        a.iteration()
        b.iteration()
    torch.cuda.synchronize()
    assert torch.equal(a.T, b.T) and torch.equal(a.F_up, b.F_up)
