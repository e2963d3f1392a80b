"""
CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A unit-free numpy + scipy restatement (fp64, CGS) of the radiative-equilibrium
hot path of bmorris3/frei.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
this module; nothing under ``frei_b200/`` does.

The reference itself cannot be imported in the build container or on the GPU
box (astropy, xarray, specutils, periodictable, pyfastchem are absent and
there is no network), so this file restates its arithmetic.  Every function
cites the reference ``file:line`` it follows (paths relative to the reference
checkout).  The (P, T) interpolation calls ``scipy.interpolate.interpn`` /
``interp1d`` — the very functions xarray's ``DataArray.interp`` dispatches to
in ``frei/opacity.py:261-263`` — so that step is the reference's own
arithmetic, not a restatement.

Pinning (``tests/test_oracle.py``): (1) every known-answer value the reference's
tests hold for this path (``frei/tests/test_core.py:42,44,52-56,60-64,67-71``);
(2) ``tests/golden/reference_run.json`` — outputs of the reference's OWN source
files (``frei/twostream.py``, ``opacity.py``, ``core.py``, ``tp.py``,
``chemistry.py``) executed here under dependency stubs for astropy.units, xarray,
specutils and periodictable (``tests/golden/run_reference.py``, ``refstubs/``):
kappa, one-step and fully converged ``emission_spectrum``, ``propagate_fluxes``,
the layer thermodynamics, the mock chemistry and the dashboard's contribution function
(``frei/plot.py`` run against a matplotlib stand-in) agree to 1e-9 .. 1e-13.

Conventions: layer index 0 = bottom (highest pressure) (``frei/tp.py:32``);
wavelength ascending.  Pressures are carried in bar where the reference does
(table axis, ``frei/opacity.py:253``) and in dyn cm^-2 for the physics;
wavelengths in micron for grids/tables and cm for the physics; fluxes in
erg s^-1 cm^-3 (``frei/twostream.py:13``).
"""
import re

import numpy as np
from scipy.interpolate import interpn, interp1d

__all__ = [
    'h', 'c', 'k_B', 'm_p', 'amu', 'sigma_sb', 'BAR',
    'pressure_grid', 'temperature_grid', 'wavelength_grid', 'BB', 'F_TOA',
    'hot_jupiter', 'rayleigh_sigma', 'bracket', 'bilinear_weights',
    'kappa', 'kappa_explicit', 'E', 'propagate_fluxes', 'bolometric_flux',
    'layer_thermo', 'emit', 'absorb', 'emission_spectrum',
    'effective_temperature', 'pressure_milne', 'contribution_function', 'load_example_opacity', 'iso_to_mass',
    'mock_mmr', 'trapz_weights', 'thermo_from_bol',
]

# ---------------------------------------------------------------------------
# Constants: CODATA 2018 / IAU 2015 values as shipped by astropy >= 4.0, in CGS
# (the reference imports them at frei/twostream.py:3, frei/core.py:3).
# ---------------------------------------------------------------------------
h = 6.62607015e-27            # erg s
c = 2.99792458e10             # cm / s
k_B = 1.380649e-16            # erg / K
m_p = 1.67262192369e-24       # g
amu = 1.66053906660e-24       # g
sigma_sb = 5.6703744191844314e-5   # erg s^-1 cm^-2 K^-4  (2 pi^5 k^4 / 15 h^3 c^2)
G = 6.6743e-8                 # cm^3 g^-1 s^-2
GM_jup = 1.2668653e23         # cm^3 s^-2
R_jup = 7.1492e9              # cm
au = 1.495978707e13           # cm
R_sun = 6.957e10              # cm
BAR = 1e6                     # dyn cm^-2 per bar

n_ref_H2 = 2.68678e19         # cm^-3, frei/opacity.py:23
n_ref_He = 2.546899e19        # cm^-3, frei/opacity.py:24
K_lambda = 1                  # frei/opacity.py:25

try:                          # numpy >= 2 spells it trapezoid (np.trapz deprecated)
    _trapz = np.trapezoid
except AttributeError:        # pragma: no cover
    _trapz = np.trapz


# ---------------------------------------------------------------------------
# Grids  (frei/tp.py, frei/core.py:34-62, 92-106)
# ---------------------------------------------------------------------------
def pressure_grid(n_layers=30, P_toa=-6, P_boa=1.1):
    """Bottom->top log-spaced pressures [bar].  frei/tp.py:10-33."""
    return np.logspace(P_toa, P_boa, n_layers)[::-1]


def temperature_grid(pressures, T_ref=2300.0, P_ref=0.1, alpha=0.1):
    """Power-law initial T-P guess [K].  frei/tp.py:36-62."""
    return T_ref * (pressures / P_ref) ** alpha


def wavelength_grid(min_micron=0.5, max_micron=10, n_bins=500, lam=None):
    """Log-spaced wavelengths [um], bin edges, resolution.  frei/core.py:34-45."""
    if lam is None:
        lam = np.logspace(np.log10(min_micron), np.log10(max_micron), n_bins)
    wl_bins = np.concatenate([[lam.min() - (lam[1] - lam[0])], lam]) + (lam[1] - lam[0]) / 2
    R = float(lam[lam.shape[0] // 2] / (lam[lam.shape[0] // 2 + 1] - lam[lam.shape[0] // 2]))
    return lam, wl_bins, R


def BB(temperature, lam_cm):
    """Planck B_lambda without the steradian, erg s^-1 cm^-3.  frei/twostream.py:46-67."""
    return (2 * h * c ** 2 / np.power(lam_cm, 5) /
            np.expm1(h * c / (lam_cm * k_B * temperature)))


def F_TOA(lam_cm, T_star=5800.0, f=2 / 3, a_rstar=0.03 * au / R_sun):
    """Stellar flux at the top of the atmosphere.  frei/core.py:48-62."""
    return f * a_rstar ** -2 * 1 / (2 * np.pi) * (np.pi * BB(T_star, lam_cm))


def hot_jupiter():
    """Planet.from_hot_jupiter(), frei/core.py:92-106, as a plain dict (CGS)."""
    return dict(a_rstar=float(0.03 * au / R_sun), m_bar=2.4 * m_p,
                g=GM_jup / R_jup ** 2, T_star=5800.0, alpha=1)


# ---------------------------------------------------------------------------
# Opacity: Rayleigh + (P, T) interpolation  (frei/opacity.py:173-269)
# ---------------------------------------------------------------------------
def n_lambda_H2(lam_cm):
    """frei/opacity.py:173-177 (Malik 2017 Eqn 17)."""
    return 13.58e-5 * (1 + 7.52e-11 * lam_cm ** -2.0) + 1


def n_lambda_He(lam_um):
    """frei/opacity.py:180-184 (Deitrick 2020 Eqn C3)."""
    return 1e-8 * (2283 + (1.8102e13 / (1.5342e10 - lam_um ** -2.0))) + 1


def rayleigh_sigma(lam_um, m_bar=2.4 * m_p):
    """sigma_H2 + sigma_He [cm^2 g^-1].  frei/opacity.py:187-200, 233."""
    lam_cm = lam_um * 1e-4
    nh2 = n_lambda_H2(lam_cm)
    nhe = n_lambda_He(lam_um)
    s_h2 = (24 * np.pi ** 3 / n_ref_H2 ** 2 / lam_cm ** 4 *
            ((nh2 ** 2 - 1) / (nh2 ** 2 + 2)) ** 2 * K_lambda) / m_bar
    s_he = (24 * np.pi ** 3 / n_ref_He ** 2 / lam_cm ** 4 *
            ((nhe ** 2 - 1) / (nhe ** 2 + 2)) ** 2 * K_lambda) / m_bar
    return s_h2 + s_he


def bracket(axis, v):
    """
    Bracket rule of scipy's ``find_indices`` (scipy >= 1.10; reached from
    frei/opacity.py:261-263 through xarray -> interpn): below the grid -> 0,
    above or on the last node -> n-2, otherwise axis[i] <= v < axis[i+1].
    Returns (i, w, out_of_bounds) with w = (v - axis[i]) / (axis[i+1] - axis[i]).
    NaN -> (-1, nan, False) like scipy.
    """
    axis = np.asarray(axis, dtype=np.float64)
    n = axis.shape[0]
    if v != v:
        return -1, np.nan, False
    oob = bool(v < axis[0] or v > axis[-1])
    if v < axis[0]:
        i = 0
    elif v >= axis[-1]:
        i = n - 2
    else:
        i = int(np.searchsorted(axis, v, side='right') - 1)
        i = min(max(i, 0), n - 2)
    denom = axis[i + 1] - axis[i]
    w = (v - axis[i]) / denom if denom != 0 else 0.0
    return i, w, oob


def bilinear_weights(wp, wt):
    """Corner weights in scipy's hypercube order (i,j),(i,j+1),(i+1,j),(i+1,j+1)."""
    return ((1.0 * (1 - wp)) * (1 - wt), (1.0 * (1 - wp)) * wt,
            (1.0 * wp) * (1 - wt), (1.0 * wp) * wt)


def _table_has_T(table):
    """frei/opacity.py:256: the temperature axis is used only if it has >1 unique value."""
    return len(np.unique(table['T'])) > 1


def kappa(tables, temperature, pressure_bar, lam_um, mmr, m_bar=2.4 * m_p):
    """
    k(lambda) = sum_s mmr_s * interp_{P,T}(table_s) + sigma;  returns (k, sigma).
    frei/opacity.py:203-269.  ``tables`` is an ordered dict/list of
    dict(P=[N_P] bar ascending, T=[N_T] K ascending, values=[N_P, N_T, n_lam]);
    ``mmr`` is the per-species mass-mixing ratio that ``chemistry()`` returns
    (frei/opacity.py:246-248) in the same order.
    """
    sigma = rayleigh_sigma(lam_um, m_bar)
    tabs = list(tables.values()) if isinstance(tables, dict) else list(tables)
    ops = []
    for s, tab in enumerate(tabs):
        if _table_has_T(tab):
            sp = _localize(tab['P'], pressure_bar)
            st = _localize(tab['T'], temperature)
            val = interpn((tab['P'][sp], tab['T'][st]), tab['values'][sp, st],
                          np.array([[pressure_bar, temperature]]),
                          method='linear', bounds_error=False, fill_value=0)[0]
        else:
            vals = tab['values'][:, 0, :] if tab['values'].ndim == 3 else tab['values']
            sp = _localize(tab['P'], pressure_bar)
            val = interp1d(tab['P'][sp], vals[sp], kind='linear', axis=0,
                           bounds_error=False, fill_value=0)(pressure_bar)
        ops.append(mmr[s] * val)
    if len(ops) == 1:
        tot = ops[0]
    else:
        tot = np.sum(np.stack(ops, axis=0), axis=0)     # xr.concat(...).sum, :268
    return tot + sigma, sigma


def _localize(axis, v):
    """
    xarray restricts the table to the neighbourhood of the requested point before
    handing it to scipy (linear/nearest methods): nearest node -2 .. +2.  The
    interpolant only touches the bracketing cell, so the value is unchanged; it
    only avoids scipy re-validating the whole table on every call.
    """
    axis = np.asarray(axis)
    i = int(np.argmin(np.abs(axis - v)))
    return slice(max(i - 2, 0), i + 3)


def kappa_explicit(tables, temperature, pressure_bar, lam_um, mmr, m_bar=2.4 * m_p):
    """Same as :func:`kappa` with the bracket/weights written out (no scipy)."""
    sigma = rayleigh_sigma(lam_um, m_bar)
    tabs = list(tables.values()) if isinstance(tables, dict) else list(tables)
    tot = np.zeros_like(lam_um, dtype=np.float64)
    for s, tab in enumerate(tabs):
        ip, wp, oob_p = bracket(tab['P'], pressure_bar)
        if _table_has_T(tab):
            it, wt, oob_t = bracket(tab['T'], temperature)
            w = bilinear_weights(wp, wt)
            v = tab['values']
            val = (0.0 + v[ip, it] * w[0] + v[ip, it + 1] * w[1] +
                   v[ip + 1, it] * w[2] + v[ip + 1, it + 1] * w[3])
            if oob_p or oob_t:
                val = np.zeros_like(val)
        else:
            v = tab['values'][:, 0, :] if tab['values'].ndim == 3 else tab['values']
            val = v[ip] + wp * (v[ip + 1] - v[ip])
            if oob_p:
                val = np.zeros_like(val)
        tot = tot + mmr[s] * val
    return tot + sigma, sigma


# ---------------------------------------------------------------------------
# Two-stream layer response  (frei/twostream.py:70-177)
# ---------------------------------------------------------------------------
def E(omega_0, g_0=0):
    """Deitrick (2020) Eqn 19.  frei/twostream.py:70-94."""
    return np.where(
        omega_0 > 0.1,
        1.225 - 0.1582 * g_0 - 0.1777 * omega_0 - 0.07465 *
        g_0 ** 2 + 0.2351 * omega_0 * g_0 - 0.05582 * omega_0 ** 2,
        np.asarray(omega_0).dtype.type(1)
    )


def propagate_fluxes(lam_cm, F_1_up, F_2_down, T_1, T_2, delta_tau, omega_0=0, g_0=0):
    """frei/twostream.py:97-177 with the same operation order."""
    wd = np.result_type(np.float64, getattr(delta_tau, 'dtype', np.float64))
    omega_0 = np.asarray(omega_0, dtype=wd).flatten()
    delta_tau = np.asarray(delta_tau, dtype=wd).flatten()
    if wd != np.float64:         # extended-precision evaluation of the same formulas (tests)
        lam_cm, F_1_up, F_2_down = (np.asarray(x, dtype=wd) for x in (lam_cm, F_1_up, F_2_down))
        T_1, T_2 = wd.type(T_1), wd.type(T_2)
    Ew = E(omega_0, g_0)

    T = np.exp(-2 * (Ew * (Ew - omega_0) * (1 - omega_0 * g_0)) ** 0.5 * delta_tau)

    zeta_plus = 0.5 * (1 + ((Ew - omega_0) / Ew / (1 - omega_0 * g_0)) ** 0.5)
    zeta_minus = 0.5 * (1 - ((Ew - omega_0) / Ew / (1 - omega_0 * g_0)) ** 0.5)

    chi = zeta_minus ** 2 * T ** 2 - zeta_plus ** 2
    xi = zeta_plus * zeta_minus * (1 - T ** 2)
    psi = (zeta_minus ** 2 - zeta_plus ** 2) * T
    pi = np.pi * (1 - omega_0) / (Ew - omega_0)

    B1 = BB(T_1, lam_cm)
    B2 = BB(T_2, lam_cm)
    Bprime = (B1 - B2) / delta_tau

    F_2_up = (
        1 / chi * (
            psi * F_1_up - xi * F_2_down +
            pi * (B2 * (chi + xi) - psi * B1 +
                  Bprime / (2 * Ew * (1 - omega_0 * g_0)) *
                  (chi - psi - xi))
        )
    )
    F_1_down = (
        1 / chi * (
            psi * F_2_down - xi * F_1_up +
            pi * (B1 * (chi + xi) - psi * B2 +
                  Bprime / (2 * Ew * (1 - omega_0 * g_0)) *
                  (xi + psi - chi))
        )
    )
    return F_2_up, F_1_down


def bolometric_flux(flux, lam_cm):
    """np.trapz over wavelength.  frei/twostream.py:16-20."""
    return _trapz(flux, lam_cm)


def trapz_weights(lam_cm):
    """w_j with sum_j w_j F_j == trapz(F, lam) (used by the sharded path)."""
    w = np.empty_like(lam_cm)
    w[1:-1] = 0.5 * (lam_cm[2:] - lam_cm[:-2])
    w[0] = 0.5 * (lam_cm[1] - lam_cm[0])
    w[-1] = 0.5 * (lam_cm[-1] - lam_cm[-2])
    return w


# ---------------------------------------------------------------------------
# Per-layer thermodynamics  (frei/twostream.py:23-43, 180-287)
# ---------------------------------------------------------------------------
def c_p(m_bar=2.4 * m_p, n_dof=5):
    """frei/twostream.py:220-224."""
    return (2 + n_dof) / (2 * m_bar) * k_B


def delta_z_i(T_i, p_i, p_ip1, g, m_bar=2.4 * m_p):
    """frei/twostream.py:180-187."""
    return (k_B * T_i) / (m_bar * g) * np.log(p_i / p_ip1)


def rho_p(p_1, p_2, T_1, g, m_bar=2.4 * m_p):
    """frei/twostream.py:234-238."""
    return ((p_1 - p_2) / g) / delta_z_i(T_1, p_1, p_2, g, m_bar)


def delta_gamma(T_i, T_ip1, p_i, p_ip1, g, m_bar=2.4 * m_p, n_dof=5):
    """frei/twostream.py:241-266."""
    return ((T_i - T_ip1) / delta_z_i(T_i, p_i, p_ip1, g, m_bar=m_bar) -
            g / c_p(m_bar=m_bar, n_dof=n_dof))


def convective_flux(T_i, T_ip1, p_i, p_ip1, g, m_bar=2.4 * m_p, n_dof=5, alpha=1):
    """frei/twostream.py:269-287."""
    rho = rho_p(p_i, p_ip1, T_i, g, m_bar=m_bar)
    cp = c_p(m_bar=m_bar, n_dof=n_dof)
    lmix = alpha * k_B * T_i / (m_bar * g)
    dg = delta_gamma(T_i, T_ip1, p_i, p_ip1, g, m_bar=m_bar, n_dof=n_dof)
    if dg > 0:
        return rho * cp * lmix ** 2 * (g / T_i) ** 0.5 * dg ** 1.5
    return 0.0


def div_bol_net_flux(F_ip1_u, F_ip1_d, F_i_u, F_i_d, T_i, T_ip1, p_i, p_ip1, g,
                     m_bar=2.4 * m_p, n_dof=5, alpha=1):
    """frei/twostream.py:190-205."""
    delta_F_rad = (F_ip1_u - F_ip1_d) - (F_i_u - F_i_d)
    delta_F_conv = convective_flux(T_i, T_ip1, p_i, p_ip1, g,
                                   m_bar=m_bar, n_dof=n_dof, alpha=alpha)
    dz = delta_z_i(T_i, p_i, p_ip1, g, m_bar)
    return (delta_F_rad + delta_F_conv) / dz, dz


def delta_t_i(p_1, p_2, T_1, T_2, delta_F_i_dz, g, m_bar=2.4 * m_p, n_dof=5):
    """frei/twostream.py:23-43 (Malik 2017 Eqns 27-28)."""
    dz = delta_z_i(T_1, p_1, p_2, g, m_bar)
    if (delta_F_i_dz * dz) != 0:
        f_i_pre = 1e5 / abs(delta_F_i_dz * dz) ** 0.9
    else:
        f_i_pre = 1
    dt_radiative = c_p(m_bar=m_bar, n_dof=n_dof) * p_1 / sigma_sb / g / T_1 ** 3
    d_gamma = delta_gamma(T_1, T_2, p_1, p_2, g, m_bar=m_bar, n_dof=n_dof)
    if d_gamma > 0:
        dt_convective = (T_1 / g / d_gamma) ** 0.5
        return f_i_pre * min(dt_radiative, dt_convective)
    return f_i_pre * dt_radiative


def delta_temperature(div, p_1, p_2, T_1, dt, g, m_bar=2.4 * m_p, n_dof=5):
    """frei/twostream.py:208-217."""
    return 1 / rho_p(p_1, p_2, T_1, g, m_bar) / c_p(m_bar, n_dof) * div * dt


def layer_thermo(Fb, T_1, T_2, p_1, p_2, g, m_bar, alpha):
    """
    dT of one layer from its four bolometric fluxes Fb = (F2_up, F2_down,
    F1_up, F1_down); the tail of the loop bodies frei/twostream.py:396-405 and
    :524-533.  ``delta_temperature`` is called without m_bar (default 2.4 m_p).
    """
    div, dz = div_bol_net_flux(Fb[0], Fb[1], Fb[2], Fb[3],
                               T_1, T_2, p_1, p_2, g, alpha=alpha, m_bar=m_bar)
    dt = delta_t_i(p_1, p_2, T_1, T_2, div, g, m_bar=m_bar)
    return delta_temperature(div, p_1, p_2, T_1, dt, g)


# ---------------------------------------------------------------------------
# Sweeps  (frei/twostream.py:290-550), one pass (n_timesteps=1 as core.py calls them)
# ---------------------------------------------------------------------------
def _bol4(F_2_up, F_2_down, F_1_up, F_1_down, lam_cm, trapz_w):
    """The four bolometric_flux calls of a layer-step (frei/twostream.py:396-398, 524-527).
    With ``trapz_w`` (weights of the global grid) a wavelength shard returns its share."""
    if trapz_w is None:
        return (bolometric_flux(F_2_up, lam_cm), bolometric_flux(F_2_down, lam_cm),
                bolometric_flux(F_1_up, lam_cm), bolometric_flux(F_1_down, lam_cm))
    return (np.dot(trapz_w, F_2_up), np.dot(trapz_w, F_2_down),
            np.dot(trapz_w, F_1_up), np.dot(trapz_w, F_1_down))


def emit(tables, temperatures, pressures_bar, lam_um, F_toa, g, m_bar, mmr_fn,
         alpha=1, fluxes_up=None, fluxes_down=None, kappa_fn=kappa, trapz_w=None,
         work_dtype=np.float64):
    """
    Upward sweep i = 1 .. L-1.  frei/twostream.py:290-421 with n_timesteps=1.
    ``mmr_fn(T, P_bar) -> [S]`` stands in for ``chemistry()``
    (frei/opacity.py:246).  Mutates fluxes_up/fluxes_down in place.
    Returns (fluxes_up, fluxes_down, T_new, T_hist[L,2], dtaus[L,n_lam], dT,
    bol[L,4]) — ``bol`` (the four np.trapz values per layer-step) is extra.
    """
    L = len(pressures_bar)
    nlam = len(lam_um)
    lam_cm = lam_um * 1e-4
    P = np.asarray(pressures_bar, dtype=np.float64) * BAR
    if fluxes_up is None:
        fluxes_up = np.zeros((L, nlam))                       # :334-335
    if fluxes_down is None:
        fluxes_down = np.zeros((L, nlam))                     # :337-339
        fluxes_down[-1] = F_toa
    temps = np.array(temperatures, dtype=np.float64)
    dtaus = [np.ones(nlam)]                                   # :352
    dT = np.zeros(L)
    bol = np.zeros((L, 4))
    for i in range(1, L):                                     # :356
        if i == L - 1:                                        # :358-363
            p_2 = P[i] * P[-2] / P[-3]
            T_2 = temps[i]
        else:
            p_2 = P[i + 1]
            T_2 = temps[i + 1]
        p_1 = P[i]
        T_1 = temps[i]
        k, sig = kappa_fn(tables, T_1, p_1 / BAR, lam_um, mmr_fn(T_1, p_1 / BAR), m_bar)
        k, sig = k.astype(work_dtype), sig.astype(work_dtype)
        delta_tau = (p_1 - p_2) / g * k                       # :227-231, :371-373
        dtaus.append(delta_tau)
        omega_0 = sig / (sig + k)                             # :376-378
        F_2_down = fluxes_down[i + 1] if i < L - 1 else F_toa  # :379-382
        F_1_up = fluxes_up[i]
        F_2_up, F_1_down = propagate_fluxes(lam_cm, F_1_up, F_2_down, T_1, T_2,
                                            delta_tau, omega_0=omega_0, g_0=0)
        Fb = _bol4(F_2_up, F_2_down, F_1_up, F_1_down, lam_cm, trapz_w)
        if i < L - 1:                                         # :392-394
            fluxes_up[i + 1] = F_2_up
        fluxes_down[i] = F_1_down
        bol[i] = Fb
        if trapz_w is None:
            dT[i] = layer_thermo(Fb, T_1, T_2, p_1, p_2, g, m_bar, alpha)
    T_new = temps - dT                                        # :407
    hist = np.stack([temps, T_new], axis=1)
    return fluxes_up, fluxes_down, T_new, hist, np.array(dtaus), dT, bol


def absorb(tables, temperatures, pressures_bar, lam_um, F_toa, g, m_bar, mmr_fn,
           alpha=1, fluxes_up=None, fluxes_down=None, kappa_fn=kappa, trapz_w=None,
           work_dtype=np.float64):
    """
    Downward sweep i = L-2 .. 0.  frei/twostream.py:424-550 with n_timesteps=1.
    ``dtaus`` rows come out in the order the loop visits them (reversed layers).
    """
    L = len(pressures_bar)
    nlam = len(lam_um)
    lam_cm = lam_um * 1e-4
    P = np.asarray(pressures_bar, dtype=np.float64) * BAR
    temps = np.array(temperatures, dtype=np.float64)
    if fluxes_up is None:                                     # :468-470
        fluxes_up = np.zeros((L, nlam))
        fluxes_up[0] = np.pi * BB(temps[0], lam_cm)
    if fluxes_down is None:                                   # :472-474
        fluxes_down = np.zeros((L, nlam))
        fluxes_down[-1] = F_toa
    dtaus = [np.ones(nlam)]                                   # :487
    dT = np.zeros(L)
    bol = np.zeros((L, 4))
    for i in range(L - 2, -1, -1):                            # :491
        p_2 = P[i + 1]
        T_2 = temps[i + 1]
        p_1 = P[i]
        T_1 = temps[i]
        k, sig = kappa_fn(tables, T_1, p_1 / BAR, lam_um, mmr_fn(T_1, p_1 / BAR), m_bar)
        k, sig = k.astype(work_dtype), sig.astype(work_dtype)
        delta_tau = (p_1 - p_2) / g * k
        dtaus.append(delta_tau)
        omega_0 = sig / (sig + k)
        F_2_down = fluxes_down[i + 1].copy()
        F_1_up = fluxes_up[i].copy()
        F_2_up, F_1_down = propagate_fluxes(lam_cm, F_1_up, F_2_down, T_1, T_2,
                                            delta_tau, omega_0=omega_0, g_0=0)
        fluxes_up[i + 1] = F_2_up                             # :521-522
        fluxes_down[i] = F_1_down
        Fb = _bol4(F_2_up, F_2_down, F_1_up, F_1_down, lam_cm, trapz_w)
        bol[i] = Fb
        if trapz_w is None:
            dT[i] = layer_thermo(Fb, T_1, T_2, p_1, p_2, g, m_bar, alpha)
    T_new = temps - dT                                        # :536
    hist = np.stack([temps, T_new], axis=1)
    return fluxes_up, fluxes_down, T_new, hist, np.array(dtaus), dT, bol


# ---------------------------------------------------------------------------
# Outer loop  (frei/core.py:233-338)
# ---------------------------------------------------------------------------
def converged_layers(temp_hists, dT, n_zero_crossings, convergence_dT):
    """frei/core.py:306-311."""
    temp_hist = np.hstack(temp_hists)
    temp_hist = temp_hist.T[temp_hist[0] != 0].T
    diffs = np.diff(temp_hist.T, axis=0)
    conv = (np.count_nonzero(np.sign(diffs[1:]) != np.sign(diffs[:-1]), axis=0)
            > n_zero_crossings) | (np.abs(dT) < convergence_dT)
    return conv, temp_hist


def emission_spectrum(tables, init_temperatures, pressures_bar, lam_um, planet, mmr_fn,
                      n_timesteps=1, n_zero_crossings=2, convergence_dT=3.0,
                      kappa_fn=kappa, return_state=False):
    """
    Grid.emission_spectrum, frei/core.py:233-338.  Returns (spectrum[n_lam] =
    fluxes_up[-1], final_temps, temp_hist, dtaus, n_iterations).
    """
    lam_cm = lam_um * 1e-4
    F_toa = F_TOA(lam_cm, T_star=planet['T_star'], a_rstar=planet['a_rstar'])  # :262
    final_temps = np.array(init_temperatures, dtype=np.float64)
    L, nlam = len(pressures_bar), len(lam_um)
    fluxes_down = np.zeros((L, nlam))                         # :265-266
    fluxes_up = np.zeros((L, nlam))
    temp_hists = []
    n_iter = 0
    for it in range(n_timesteps):                             # :273
        fluxes_up, fluxes_down, final_temps, _, _, dT, _ = emit(
            tables, final_temps, pressures_bar, lam_um, F_toa, planet['g'],
            planet['m_bar'], mmr_fn, alpha=planet['alpha'],
            fluxes_up=fluxes_up, fluxes_down=fluxes_down, kappa_fn=kappa_fn)
        fluxes_up, fluxes_down, final_temps, hist_absorb, _, dT, _ = absorb(
            tables, final_temps, pressures_bar, lam_um, F_toa, planet['g'],
            planet['m_bar'], mmr_fn, alpha=planet['alpha'],
            fluxes_up=fluxes_up, fluxes_down=fluxes_down, kappa_fn=kappa_fn)
        n_iter += 1
        temp_hists.append(hist_absorb)
        conv, _ = converged_layers(temp_hists, dT, n_zero_crossings, convergence_dT)
        if np.all(conv):                                      # :317
            break
    temp_hist = np.hstack(temp_hists)                         # :320-321
    temp_hist = temp_hist.T[temp_hist[0] != 0].T
    # final emit: alpha is NOT forwarded (default 1), frei/core.py:323-333
    fluxes_up, fluxes_down, final_temps, _, dtaus, dT, _ = emit(
        tables, final_temps, pressures_bar, lam_um, F_toa, planet['g'],
        planet['m_bar'], mmr_fn, alpha=1,
        fluxes_up=fluxes_up, fluxes_down=fluxes_down, kappa_fn=kappa_fn)
    out = (fluxes_up[-1].copy(), final_temps, temp_hist, dtaus, n_iter)
    if return_state:
        out = out + (fluxes_up, fluxes_down)
    return out


# ---------------------------------------------------------------------------
# T_eff diagnostics  (frei/core.py:386-439) — used by the reference's KAT
# ---------------------------------------------------------------------------
def pressure_milne(pressures_bar, dtaus):
    """Per-wavelength tau ~ 2/3 pressure: the loop of frei/core.py:390-395 (np.interp itself)."""
    out = np.ones(dtaus.shape[1])
    for i in range(dtaus.shape[1]):
        out[i] = np.interp(2 / 3, np.exp(-dtaus[:, i]), pressures_bar)
    return out


def effective_temperature_milne(pressures_bar, lam_um, spec, dtaus, final_temps):
    """frei/core.py:386-405."""
    lam_cm = lam_um * 1e-4
    return np.interp(np.average(pressure_milne(pressures_bar, dtaus), weights=spec * lam_cm),
                     pressures_bar[::-1], final_temps[::-1])


def contribution_function(lam_um, pressures_bar, temps, dtaus):
    """
    Normalised contribution function of the dashboard, frei/plot.py:63-79, returned as the
    reference plots it (``cf[::-1]``, level order, frei/plot.py:83).  The reference has no test or
    fixture for it; pinned to the array its own ``dashboard`` hands to ``pcolormesh`` when run
    under a matplotlib stand-in (tests/golden/run_reference.py, case E).
    """
    dtaus = np.asarray(dtaus)
    tau = np.cumsum(dtaus[::-1], axis=0)                                   # :63
    nus = 1.0 / (lam_um * 1e-4)                                            # :64, lam -> cm^-1
    hcperk = h * c / k_B                                                   # :65
    dlogP = (np.log10(pressures_bar.max()) - np.log10(pressures_bar.min())) / (len(pressures_bar) - 1)
    k = 10 ** -dlogP                                                       # :70
    dParr = (1 - k) * pressures_bar                                        # :71
    cf = (np.exp(-tau) * dtaus[::-1] * (pressures_bar[::-1, None] / dParr[::-1, None]) *
          nus ** 3 / np.expm1(hcperk * nus / temps[::-1, None]))           # :73-77
    cf /= np.sum(cf, axis=0)                                               # :79
    return cf[::-1]


def effective_temperature_planck(lam_um, spec):
    """frei/core.py:408-414."""
    bol_flux = _trapz(spec, lam_um * 1e-4)
    return (bol_flux / sigma_sb) ** (1 / 4)


def effective_temperature(pressures_bar, lam_um, spec, dtaus, final_temps):
    """Mean of the Milne and Stefan-Boltzmann estimates.  frei/core.py:417-439."""
    return np.mean([effective_temperature_milne(pressures_bar, lam_um, spec, dtaus, final_temps),
                    effective_temperature_planck(lam_um, spec)])


# ---------------------------------------------------------------------------
# Fixture + mock chemistry  (frei/opacity.py:272-342, frei/chemistry.py:24-37, 197-246)
# ---------------------------------------------------------------------------
def example_band_profile(lam_um, seed=42):
    """The ``so`` curve of load_example_opacity, frei/opacity.py:295-324."""
    np.random.seed(seed)
    so = (np.exp(-0.5 * (lam_um - 6) ** 2 / 2 ** 2) +
          0.8 * np.exp(-0.5 * (lam_um - 0.3) ** 2 / 0.5 ** 2))
    amps = np.random.uniform(low=0.1, high=0.2, size=15)
    cens = np.random.uniform(low=0.5, high=1, size=15)
    for amp, wl in zip(amps, cens):
        so = so + amp * np.exp(-0.5 * (lam_um - wl) ** 2 / 0.005 ** 2)
    for amp, wl in zip([0.22, 0.2, 0.18], np.logspace(np.log10(1.4), np.log10(2.7), 3)):
        so = so + amp * np.exp(-0.5 * (lam_um - wl) ** 2 / 0.13 ** 2)
    return so


def load_example_opacity(pressures_bar, init_temperatures, lam_um, seed=42, scale_factor=20):
    """
    frei/opacity.py:272-342 as {"1H2-16O": table}.  The reference stores the
    axes in Grid order (descending) and xarray sorts them on interp; here they
    are returned ascending.  ``drop_duplicates('temperature')`` (:339) keeps the
    first occurrence.
    """
    so = example_band_profile(lam_um, seed)
    simple = np.zeros((len(pressures_bar), len(init_temperatures), len(lam_um)))
    simple[:] += 5 * 10 ** (2.5 * (so - 0.4))
    simple *= scale_factor
    P = np.asarray(pressures_bar, dtype=np.float64)
    T = np.asarray(init_temperatures, dtype=np.float64)
    _, first = np.unique(T, return_index=True)
    keep = np.sort(first)
    T = T[keep]
    simple = simple[:, keep, :]
    ip = np.argsort(P, kind='stable')
    it = np.argsort(T, kind='stable')
    return {"1H2-16O": dict(P=P[ip], T=T[it], values=np.ascontiguousarray(simple[ip][:, it]))}


def iso_to_mass(isotopologue):
    """'1H2-16O' -> 18 (in u).  frei/chemistry.py:24-37 (numeric branch only)."""
    mass = 0
    for element in isotopologue.split('-'):
        multiples = list(filter(lambda x: len(x) > 0, re.split(r'\D', element)))
        if len(multiples) > 1:
            species_mass, multiplier = multiples
            mass += float(multiplier) * float(species_mass)
        elif len(multiples) == 1:
            mass += float(multiples[0])
    return mass


def mock_mmr(species, m_bar=2.4 * m_p, vmr=1.5e-3):
    """
    Mass-mixing ratios when pyfastchem is absent: VMR = 1.5e-3 for every species
    (frei/chemistry.py:232-246), mmr = vmr * mass / m_bar (frei/chemistry.py:197-199).
    """
    return np.array([vmr * (iso_to_mass(s) * amu / m_bar) for s in species])


def thermo_from_bol(bol, temperatures, pressures_bar, g, m_bar, alpha, direction):
    """
    dT[L] of one sweep from the (globally summed) four integrals per layer-step:
    the scalar tail of the loop bodies (frei/twostream.py:396-407, 524-536).
    ``direction`` 'emit' | 'absorb'.
    """
    P = np.asarray(pressures_bar, dtype=np.float64) * BAR
    T = np.asarray(temperatures, dtype=np.float64)
    L = len(P)
    dT = np.zeros(L)
    rng = range(1, L) if direction == 'emit' else range(L - 2, -1, -1)
    for i in rng:
        if i == L - 1:
            p_2, T_2 = P[i] * P[-2] / P[-3], T[i]
        else:
            p_2, T_2 = P[i + 1], T[i + 1]
        dT[i] = layer_thermo(tuple(bol[i]), T[i], T_2, P[i], p_2, g, m_bar, alpha)
    return dT
