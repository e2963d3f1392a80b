"""
Two-stream sweeps — same interface as ``frei/twostream.py`` (``emit``,
``absorb``, ``propagate_fluxes``).  The arithmetic runs in the CUDA kernels of
``csrc/frei_b200.cu`` through the C ABI; host arrays are copied to the GPU and
back on every call, exactly as a drop-in for the reference functions must.
For a device-resident solve use :meth:`frei_b200.core.Grid.emission_spectrum`
or :class:`frei_b200.engine.Engine` directly.
"""
import numpy as np

from . import _cabi
from . import units as U
from .chemistry import chemistry
from .engine import Engine, FREI_EMIT, FREI_ABSORB
from .opacity import device_table

__all__ = ['propagate_fluxes', 'emit', 'absorb']


def propagate_fluxes(lam, F_1_up, F_2_down, T_1, T_2, delta_tau, omega_0=0, g_0=0, eps=0.5):
    """
    Fluxes leaving one layer (improved two-stream; Malik et al. 2017 Eq. 15,
    Deitrick et al. 2020 B2, 2022 B4); signature of frei/twostream.py:97-99.
    Returns (F_2_up, F_1_down).  ``g_0`` must be 0 (the only value the
    reference's callers use, frei/twostream.py:389, 518); ``eps`` is unused there too.
    """
    import torch
    if np.any(np.asarray(g_0) != 0):
        raise NotImplementedError('g_0 != 0 is not on the hot path (frei/twostream.py:389)')
    lib = _cabi.load()
    _cabi.require_cuda()
    lam_cm = np.ascontiguousarray(U.value(lam, 'cm') if U.is_quantity(lam)
                                  else np.asarray(lam, dtype=np.float64) * 1e-4)
    n = lam_cm.shape[0]
    dev = torch.device('cuda', torch.cuda.current_device())

    def up(x, unit):
        a = np.ascontiguousarray(np.broadcast_to(U.value(x, unit).flatten(), (n,)))
        return torch.from_numpy(a.copy()).to(dev)
    d_lam = torch.from_numpy(lam_cm).to(dev)
    d_f1, d_f2 = up(F_1_up, 'flux'), up(F_2_down, 'flux')
    d_tau, d_w0 = up(delta_tau, ''), up(omega_0, '')
    o1, o2 = torch.empty_like(d_f1), torch.empty_like(d_f1)
    _cabi.check(lib.frei_b200_propagate(
        d_lam.data_ptr(), d_f1.data_ptr(), d_f2.data_ptr(), float(U.value(T_1, 'K')),
        float(U.value(T_2, 'K')), d_tau.data_ptr(), d_w0.data_ptr(), o1.data_ptr(),
        o2.data_ptr(), n, torch.cuda.current_stream(dev).cuda_stream))
    return U.wrap(o1.cpu().numpy(), 'flux'), U.wrap(o2.cpu().numpy(), 'flux')


def _sweep_api(direction, opacities, temperatures, pressures, lam, F_TOA, g, m_bar,
               n_timesteps, convergence_thresh, alpha, fluxes_up, fluxes_down):
    T0 = U.value(temperatures, 'K').copy()
    P = U.value(pressures, 'bar')
    lam_um = U.value(lam, 'um')
    f_toa = U.value(F_TOA, 'flux')
    g_cgs = U.gravity_cgs(g)
    m_bar_g = float(U.value(m_bar, 'g'))
    thresh = float(U.value(convergence_thresh, 'K'))
    L, n_lam = P.shape[0], lam_um.shape[0]
    species = list(opacities.keys())

    def mmr_of(T):
        d = chemistry(T, P, species, m_bar=m_bar_g)
        return np.stack([np.broadcast_to(d[s], T.shape) for s in species], axis=-1)

    eng = Engine(device_table(opacities), lam_um, P, T0, mmr_of(T0), g=g_cgs, m_bar=m_bar_g,
                 alpha=alpha, f_toa=f_toa, want_dtaus=True)
    # default initial fluxes, frei/twostream.py:334-339 and :468-474
    up_host = None if fluxes_up is None else U.value(fluxes_up, 'flux')
    down_host = None if fluxes_down is None else U.value(fluxes_down, 'flux')
    if up_host is None:
        up_host = np.zeros((L, n_lam))
        if direction == FREI_ABSORB:
            from .core import BB
            up_host[0] = np.pi * U.value(BB(T0[0])(lam_um), 'flux')
    if down_host is None:
        down_host = np.zeros((L, n_lam))
        down_host[-1] = f_toa
    eng.set_fluxes(up_host, down_host)

    hist = np.zeros((L, n_timesteps + 1))
    hist[:, 0] = T0
    dT = np.zeros(L)
    j = 0
    for j in range(n_timesteps):
        if j > 0:
            eng.set_mmr(mmr_of(hist[:, j]))
        eng.sweep(direction, alpha_override=-1.0, with_dtaus=True)
        hist[:, j + 1] = eng.T[0].cpu().numpy()
        dT = eng.dT[0].cpu().numpy()
        if n_timesteps > 1 and np.abs(dT).max() < thresh:       # :408-416, :537-545
            break
    F_up = eng.F_up[0].cpu().numpy()
    F_down = eng.F_down[0].cpu().numpy()
    dtaus = eng.dtaus[0].cpu().numpy()
    # the reference mutates the caller's arrays in place (:392-394, :521-522)
    for dst, src in ((fluxes_up, F_up), (fluxes_down, F_down)):
        if dst is not None:
            try:
                if U.is_quantity(dst):
                    dst[...] = U.wrap(src, 'flux')
                else:
                    np.asarray(dst)[...] = src
            except (ValueError, TypeError):
                pass
    return (U.wrap(F_up, 'flux') if fluxes_up is None else fluxes_up,
            U.wrap(F_down, 'flux') if fluxes_down is None else fluxes_down,
            U.wrap(hist[:, j + 1].copy(), 'K'), U.wrap(hist, 'K'), dtaus, U.wrap(dT, 'K'))


def emit(opacities, temperatures, pressures, lam, F_TOA, g, m_bar=2.4 * U.m_p,
         n_timesteps=50, convergence_thresh=10, alpha=1, fluxes_up=None, fluxes_down=None):
    """
    Upward sweep (bottom -> top) with the temperature update; signature and
    6-tuple return ``(fluxes_up, fluxes_down, T_final, temperature_history,
    dtaus, dT)`` of frei/twostream.py:290-421.
    """
    return _sweep_api(FREI_EMIT, opacities, temperatures, pressures, lam, F_TOA, g, m_bar,
                      n_timesteps, convergence_thresh, alpha, fluxes_up, fluxes_down)


def absorb(opacities, temperatures, pressures, lam, F_TOA, g, m_bar=2.4 * U.m_p,
           n_timesteps=50, convergence_thresh=10, alpha=1, fluxes_up=None, fluxes_down=None):
    """Downward sweep (top -> bottom); signature and return of frei/twostream.py:424-550."""
    return _sweep_api(FREI_ABSORB, opacities, temperatures, pressures, lam, F_TOA, g, m_bar,
                      n_timesteps, convergence_thresh, alpha, fluxes_up, fluxes_down)
