"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY (see frei_oracle.py).

The reference's two-stream expressions (frei/twostream.py:139-176) evaluated
with mpmath at 40 significant digits: the exact value of the reference's
formulas, free of the rounding noise that their grouping picks up in fp64 (up
to ~4e-6 relative for delta_tau < 1e-6) and even in 80-bit long double (~2e-9).
Used on a handful of wavelength columns to arbitrate when the fp64 oracle and
the GPU differ by more than the fp64 oracle can resolve.
"""
import mpmath as mp
import numpy as np

from . import frei_oracle as O

mp.mp.dps = 40


def _E(w0):
    """frei/twostream.py:89-94 with g_0 = 0."""
    if w0 > mp.mpf('0.1'):
        return mp.mpf(1.225) - mp.mpf(0.1777) * w0 - mp.mpf(0.05582) * w0 ** 2
    return mp.mpf(1)


def _BB(T, lam_cm):
    """frei/twostream.py:64-67."""
    h, c, k = mp.mpf(O.h), mp.mpf(O.c), mp.mpf(O.k_B)
    return 2 * h * c ** 2 / lam_cm ** 5 / mp.expm1(h * c / (lam_cm * k * T))


def propagate_one(lam_cm, F1u, F2d, T1, T2, dtau, w0):
    """One wavelength of propagate_fluxes (frei/twostream.py:97-177), g_0 = 0."""
    lam_cm, F1u, F2d, T1, T2, dtau, w0 = (mp.mpf(float(x)) if not isinstance(x, mp.mpf) else x
                                          for x in (lam_cm, F1u, F2d, T1, T2, dtau, w0))
    Ew = _E(w0)
    T = mp.exp(-2 * mp.sqrt(Ew * (Ew - w0)) * dtau)
    r = mp.sqrt((Ew - w0) / Ew)
    zp, zm = (1 + r) / 2, (1 - r) / 2
    chi = zm ** 2 * T ** 2 - zp ** 2
    xi = zp * zm * (1 - T ** 2)
    psi = (zm ** 2 - zp ** 2) * T
    pi = mp.mpf(float(np.pi)) * (1 - w0) / (Ew - w0)
    B1, B2 = _BB(T1, lam_cm), _BB(T2, lam_cm)
    Bp = (B1 - B2) / dtau
    F2u = (psi * F1u - xi * F2d + pi * (B2 * (chi + xi) - psi * B1 + Bp / (2 * Ew) * (chi - psi - xi))) / chi
    F1d = (psi * F2d - xi * F1u + pi * (B1 * (chi + xi) - psi * B2 + Bp / (2 * Ew) * (xi + psi - chi))) / chi
    return F2u, F1d


def sweep_column(direction, k, sig, dpg, T, lam_cm, F_toa, Fu_col, Fd_col):
    """
    One wavelength column of emit / absorb (frei/twostream.py:356-394, 491-522).
    k[L], sig: total opacity per level (fp64 values taken as exact) and sigma;
    dpg[L] = (p1 - p2)/g per level; T[L]; Fu_col, Fd_col: lists of mpf, mutated.
    """
    L = len(T)
    lam_cm = mp.mpf(float(lam_cm))
    sig = mp.mpf(float(sig))
    rng = range(1, L) if direction == 'emit' else range(L - 2, -1, -1)
    for i in rng:
        ki = mp.mpf(float(k[i]))
        dtau = mp.mpf(float(dpg[i])) * ki
        w0 = sig / (sig + ki)
        top = direction == 'emit' and i == L - 1
        T2 = T[i] if top else T[i + 1]
        F2d = mp.mpf(float(F_toa)) if top else Fd_col[i + 1]
        F2u, F1d = propagate_one(lam_cm, Fu_col[i], F2d, mp.mpf(float(T[i])), mp.mpf(float(T2)),
                                 dtau, w0)
        if not top:
            Fu_col[i + 1] = F2u
        Fd_col[i] = F1d
    return Fu_col, Fd_col
