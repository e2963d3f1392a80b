"""The C-ABI library loads and exports exactly what include/frei_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

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
