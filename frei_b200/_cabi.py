"""
ctypes binding of ``include/frei_b200.h``.

The library is loaded from ``frei_b200/_lib/libfrei_b200.so`` (built in-tree by
``frei_b200/build.py``).  There is no CPU fallback: if the library is missing
and cannot be built, or a compute call is made without a CUDA device, an
exception is raised.
"""
import ctypes as C
import os

from . import build as _build

ABI_VERSION = 2          # FREI_B200_ABI_VERSION of include/frei_b200.h
FREI_EMIT, FREI_ABSORB = 0, 1
FREI_F32, FREI_F64 = 32, 64

c_void_p, c_int32, c_int64, c_double = C.c_void_p, C.c_int32, C.c_int64, C.c_double


class FreiError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""


class frei_table(C.Structure):
    _fields_ = [('values', c_void_p), ('axis_P', c_void_p), ('axis_T', c_void_p),
                ('has_T', c_void_p), ('S', c_int32), ('N_P', c_int32), ('N_T', c_int32),
                ('dtype', c_int32), ('n_lam', c_int64)]


class frei_spectral(C.Structure):
    _fields_ = [('c1', c_void_p), ('c2', c_void_p), ('sigma', c_void_p), ('w', c_void_p),
                ('f_toa', c_void_p), ('n_lam', c_int64)]


class frei_tracker(C.Structure):
    _fields_ = [('last_T', c_void_p), ('state', c_void_p), ('n_columns', c_void_p),
                ('iterations', c_void_p), ('n_zero_crossings', c_int32),
                ('convergence_dT', c_double)]


class frei_atmosphere(C.Structure):
    _fields_ = [('T', c_void_p), ('P', c_void_p), ('mmr', c_void_p), ('g', c_void_p),
                ('m_bar', c_void_p), ('alpha', c_void_p), ('sigma_scale', c_void_p),
                ('ftoa_scale', c_void_p), ('B', c_int32), ('L', c_int32),
                ('active', c_void_p), ('tracker', C.POINTER(frei_tracker))]


class frei_flux(C.Structure):
    _fields_ = [('F_up', c_void_p), ('F_down', c_void_p), ('dtaus', c_void_p),
                ('dtype', c_int32)]


class frei_p2p(C.Structure):
    _fields_ = [('peer_bufs', c_void_p), ('error', c_void_p),
                ('epoch', C.c_uint64), ('rank', c_int32), ('world', c_int32)]


class frei_workspace(C.Structure):
    _fields_ = [('layer_params', c_void_p), ('partials', c_void_p), ('sums', c_void_p),
                ('dT', c_void_p)]


P = C.POINTER
# name -> (restype, argtypes); every symbol declared in include/frei_b200.h
SIGNATURES = {
    'frei_b200_last_error': (C.c_char_p, []),
    'frei_b200_abi_version': (C.c_int, []),
    'frei_b200_device_count': (C.c_int, []),
    'frei_b200_workspace_bytes': (C.c_int, [c_int32, c_int32, c_int32, c_int64,
                                            P(c_int64), P(c_int64), P(c_int64), P(c_int64)]),
    'frei_b200_spectral_setup': (C.c_int, [c_void_p, c_int64, c_int64, c_int64, c_double, c_double,
                                           c_double, c_double, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p]),
    'frei_b200_debug_math': (C.c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'frei_b200_debug_plan': (C.c_int, [c_int32]),
    'frei_b200_debug_plan_query': (C.c_int, [c_int64, c_int32, c_int32, c_int64, c_int32, C.POINTER(c_int32)]),
    'frei_b200_layer_prep': (C.c_int, [P(frei_table), P(frei_atmosphere), P(frei_workspace),
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'frei_b200_kappa': (C.c_int, [P(frei_table), P(frei_spectral), P(frei_atmosphere),
                                  P(frei_workspace), c_void_p, c_void_p, c_void_p]),
    'frei_b200_propagate': (C.c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'frei_b200_sweep': (C.c_int, [P(frei_table), P(frei_spectral), P(frei_atmosphere),
                                  P(frei_flux), c_int32, P(frei_workspace), c_void_p]),
    'frei_b200_reduce': (C.c_int, [P(frei_atmosphere), P(frei_workspace), c_int64, c_void_p]),
    'frei_b200_update_T': (C.c_int, [P(frei_table), P(frei_atmosphere), P(frei_workspace), c_int32,
                                     c_double, c_void_p, c_void_p]),
    'frei_b200_post': (C.c_int, [P(frei_table), P(frei_atmosphere), P(frei_workspace), c_int64,
                                 c_int32, c_double, c_void_p, c_int32, c_void_p]),
    'frei_b200_post_p2p': (C.c_int, [P(frei_table), P(frei_atmosphere), P(frei_workspace), c_int64,
                                     c_int32, c_double, c_void_p, c_int32, P(frei_p2p), c_void_p]),
    'frei_b200_sweep_step': (C.c_int, [P(frei_table), P(frei_spectral), P(frei_atmosphere),
                                       P(frei_flux), c_int32, c_double, P(frei_workspace),
                                       c_void_p, c_int32, c_int32, c_void_p]),
    'frei_b200_diagnostics_scratch_bytes': (c_int64, [c_int64]),
    'frei_b200_diagnostics': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    'frei_b200_bin_trapz': (C.c_int, [c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                      c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    'frei_b200_regrid': (C.c_int, [c_void_p, c_int32, c_int32, c_int64, c_void_p, c_int32, c_void_p,
                                   c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    'frei_b200_fp64_peak': (C.c_int, [c_void_p, c_int64, P(c_double), c_void_p]),
}

_lib = None


def lib_path():
    return _build.lib_path()


def load():
    """Load (building first if the sources are newer and nvcc is present) the C-ABI library."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if os.environ.get('FREI_B200_LIB'):          # experiment builds (build.build_variant)
        path = os.environ['FREI_B200_LIB']
        if not os.path.exists(path):
            raise FreiError(f'FREI_B200_LIB={path} does not exist')
    elif _build.find_nvcc() is not None:
        path = _build.build()
    elif not os.path.exists(path):
        raise FreiError(
            f'{path} is missing and nvcc is not available to build it; '
            'frei_b200 has no CPU fallback')
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.frei_b200_abi_version() != ABI_VERSION:
        raise FreiError('libfrei_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().frei_b200_last_error().decode(errors='replace')
        raise FreiError(f'frei_b200 error {rc}: {msg}')


def require_cuda():
    """The product path needs a GPU; fail loudly otherwise."""
    import torch
    if not torch.cuda.is_available() or load().frei_b200_device_count() < 1:
        raise FreiError('no CUDA device: frei_b200 runs on the GPU only (no CPU fallback)')


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
