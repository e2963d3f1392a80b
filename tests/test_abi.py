"""The C-ABI library loads and exports exactly what include/frei_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from frei_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'frei_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(frei_b200_\w+)\s*\(', text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _cabi.load()
    names = declared_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
        assert name in _cabi.SIGNATURES, f'{name} has no ctypes signature'
    assert sorted(_cabi.SIGNATURES) == names


def test_struct_layouts_match_header():
    assert C.sizeof(_cabi.frei_table) == 4 * 8 + 4 * 4 + 8
    assert C.sizeof(_cabi.frei_spectral) == 5 * 8 + 8
    assert C.sizeof(_cabi.frei_atmosphere) == 8 * 8 + 2 * 4 + 2 * 8
    assert C.sizeof(_cabi.frei_tracker) == 4 * 8 + 4 + 4 + 8
    assert C.sizeof(_cabi.frei_flux) == 3 * 8 + 8
    assert C.sizeof(_cabi.frei_workspace) == 4 * 8


def test_abi_version_and_workspace_query():
    lib = _cabi.load()
    assert lib.frei_b200_abi_version() == 2
    sizes = [C.c_int64() for _ in range(4)]
    rc = lib.frei_b200_workspace_bytes(2, 50, 3, 200000, *[C.byref(s) for s in sizes])
    assert rc == 0
    lp, partials, sums, dT = (s.value for s in sizes)
    assert lp == 2 * 50 * 18 * 8                  # record = 2 + 5 S words, padded to even
    assert sums == 2 * 50 * 4 * 8 and dT == 2 * 50 * 8
    assert partials >= 2 * (200000 // 32) * 50 * 4 * 8
    assert lib.frei_b200_workspace_bytes(1, 2, 3, 10, None, None, None, None) == -1   # L < 3
    assert b'bad argument' in lib.frei_b200_last_error()


def test_bad_arguments_are_rejected_without_touching_the_device():
    lib = _cabi.load()
    assert lib.frei_b200_propagate(None, None, None, 1.0, 1.0, None, None, None, None, 10, None) == -1
    assert lib.frei_b200_last_error()
    assert lib.frei_b200_debug_math(None, None, 0, None) == -1


def test_diagnostics_and_plan_hooks_validate_arguments():
    lib = _cabi.load()
    assert lib.frei_b200_diagnostics_scratch_bytes(0) == 0
    assert lib.frei_b200_diagnostics_scratch_bytes(1) == 3 * 8
    assert lib.frei_b200_diagnostics_scratch_bytes(200_000) == ((200_000 + 255) // 256) * 3 * 8
    assert lib.frei_b200_diagnostics(None, None, None, None, None, None, 30, 500, None, None, None, None,
                                     None) == -1
    assert b'frei_b200_diagnostics' in lib.frei_b200_last_error()
    assert lib.frei_b200_debug_plan(5) == -1
    assert lib.frei_b200_debug_plan(3) == 0 and lib.frei_b200_debug_plan(0) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product path refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import frei_b200 as frei
    with pytest.raises(_cabi.FreiError):
        _cabi.require_cuda()
    planet = frei.Planet.from_hot_jupiter()
    grid = frei.Grid(planet, n_wl_bins=40, n_layers=8)
    grid.load_opacities(opacities=frei.load_example_opacity(grid))
    with pytest.raises(_cabi.FreiError):
        grid.emission_spectrum(n_timesteps=1)
    with pytest.raises(_cabi.FreiError):
        frei.kappa(grid.opacities, 1000.0, 1.0, grid.lam)


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under frei_b200/ may import it."""
    pkg = os.path.join(ROOT, 'frei_b200')
    pat = re.compile(r'^\s*(from|import)\s+(\.*\s*)?oracle\b|__import__\([\'"]oracle', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f'{f} imports the oracle'


def _relay_pieces(rows, NS, quota, n_warps):
    """The work items of every warp of a relay launch, in execution order — a restatement of the loop
    in sweep_kernel<..., RELAY = true>: warp m owns steps [m q, (m + 1) q) of the chunk-major step line
    and works through them backwards."""
    W = rows * NS
    out = []
    for m in range(n_warps):
        p0 = min(m, W // quota + 1) * quota
        p = min(p0 + quota, W)
        items = []
        while p > p0:
            q = (p - 1) // NS
            base = q * NS
            lo, hi = max(p0, base) - base, p - base
            p = base + lo
            items.append((q, lo, hi))
        out.append(items)
    return out


@pytest.mark.parametrize('n_lam,L,warps', [(200_000, 50, 2368), (160_000, 50, 2368), (125_000, 100, 1184),
                                            (250_000, 200, 2368), (1_000_000, 100, 1184), (151_554, 50, 2368),
                                            (303_106, 30, 2368)])
def test_relay_plan_schedule_properties(n_lam, L, warps):
    """Host logic of the relay plan (no GPU): every (chunk, layer-step) is computed exactly once, a
    chunk is cut into at most two pieces, the first piece is the first thing its warp does and belongs
    to the warp just below the one that finishes the chunk last of all, with quota - (L - 1) >= 0
    steps of slack, and no warp works longer than the quota."""
    lib = _cabi.load()
    out = (C.c_int32 * 4)()
    assert lib.frei_b200_debug_plan(0) == 0
    assert lib.frei_b200_debug_plan_query(n_lam, 1, L, warps, 1, out) == 0
    n2, n1, quota, n_warps = list(out)
    NS = L - 1
    rows = (n_lam + 63) // 64
    assert rows > warps, 'test cases are meant to need a relay plan'
    assert (n2, n1) == (rows, 0) and quota >= NS and 0 < n_warps <= warps
    assert quota == -(-rows * NS // warps)
    sched = _relay_pieces(rows, NS, quota, 4 * ((n_warps + 3) // 4))       # whole CTAs are launched
    seen = np.zeros((rows, NS), dtype=np.int32)
    pieces = {}
    for m, items in enumerate(sched):
        assert sum(hi - lo for _, lo, hi in items) <= quota
        done = 0
        for order, (q, lo, hi) in enumerate(items):
            assert 0 <= lo < hi <= NS
            seen[q, lo:hi] += 1
            pieces.setdefault(q, []).append((lo, hi, m, order, done, len(items)))
            done += hi - lo
    assert (seen == 1).all()
    for q, ps in pieces.items():
        assert len(ps) <= 2
        if len(ps) == 2:
            (lo_a, hi_a, m_a, ord_a, start_a, _), (lo_b, hi_b, m_b, ord_b, start_b, n_b) = sorted(ps)
            assert lo_a == 0 and hi_a == lo_b and hi_b == NS
            assert m_b == m_a + 1                    # the receiver is the next warp: same or next CTA
            assert ord_a == 0 and start_a == 0       # the sender runs its piece first of all
            assert ord_b == n_b - 1                  # the receiver runs its piece last
            if m_b < n_warps - 1:                    # slack in layer-steps between hand-over and use
                assert start_b - hi_a == quota - NS
            # the last warp's run may be short: it then waits for its sender, but never beyond the quota
            assert max(start_b, hi_a) + (NS - lo_b) <= quota


def test_plan_query_without_relay_and_small_problems():
    lib = _cabi.load()
    out = (C.c_int32 * 4)()
    # more resident warps than chunks: whole 64-wide chunks
    assert lib.frei_b200_debug_plan_query(100_000, 1, 50, 2368, 1, out) == 0
    assert list(out) == [1563, 0, 0, 0]
    # relay not allowed (dtaus sweep) or a batch: whole chunks as well
    assert lib.frei_b200_debug_plan_query(200_000, 1, 50, 2368, 0, out) == 0 and list(out) == [3125, 0, 0, 0]
    assert lib.frei_b200_debug_plan_query(200_000, 4, 50, 0, 1, out) == 0 and list(out) == [3125, 0, 0, 0]
    # odd wavelength count and C1 (too few chunks for the SMs): 32-wide chunks
    assert lib.frei_b200_debug_plan_query(5_001, 1, 50, 2368, 1, out) == 0 and list(out) == [0, 157, 0, 0]
    assert lib.frei_b200_debug_plan_query(5_000, 1, 50, 2368, 1, out) == 0 and list(out) == [0, 157, 0, 0]
    assert lib.frei_b200_debug_plan_query(0, 1, 50, 2368, 1, out) == -1
