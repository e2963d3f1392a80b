"""
User-facing driver — same interface as ``frei/core.py`` (``Planet``, ``Grid``,
``effective_temperature``).  ``Grid.emission_spectrum`` keeps the reference's
host control flow (outer loop, convergence test, final emit,
frei/core.py:263-338) and runs every sweep on the GPU with the flux state,
tables and per-wavelength constants resident in HBM; only T and dT
(``n_layers`` doubles each) cross PCIe per iteration.
"""
import collections

import numpy as np

from . import _cabi
from . import units as U
from .chemistry import chemistry
from .engine import Engine, DeviceTable, FREI_EMIT, FREI_ABSORB, FREI_F64, shard_range
from .tp import pressure_grid, temperature_grid

__all__ = ['Grid', 'Planet', 'effective_temperature', 'contribution_function', 'Spectrum']


def wavelength_grid(min_micron=0.5, max_micron=10, n_bins=500, lam=None):
    """Log-spaced wavelength grid, bin edges and resolution (frei/core.py:34-45)."""
    if lam is None:
        lam_um = np.logspace(np.log10(min_micron), np.log10(max_micron), n_bins)
    else:
        lam_um = U.value(lam, 'um')
    wl_bins = np.concatenate([[lam_um.min() - (lam_um[1] - lam_um[0])], lam_um]) \
        + (lam_um[1] - lam_um[0]) / 2
    mid = lam_um.shape[0] // 2
    R = float(lam_um[mid] / (lam_um[mid + 1] - lam_um[mid]))
    return U.wrap(lam_um, 'um'), wl_bins, R


def _spectral_host(lam_um, m_bar_g, T_star, a_rstar, f=2 / 3):
    """Run the spectral_setup kernel and return its five arrays on the host."""
    import torch
    lib = _cabi.load()
    _cabi.require_cuda()
    dev = torch.device('cuda', torch.cuda.current_device())
    lam_um = np.ascontiguousarray(lam_um, dtype=np.float64)
    n = lam_um.shape[0]
    d_lam = torch.from_numpy(lam_um).to(dev)
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(5)]
    _cabi.check(lib.frei_b200_spectral_setup(
        d_lam.data_ptr(), n, 0, n, float(m_bar_g), float(T_star), float(a_rstar), float(f),
        *[o.data_ptr() for o in outs], torch.cuda.current_stream(dev).cuda_stream))
    return [o.cpu().numpy() for o in outs]


def BB(temperature):
    """Planck function factory, B_lambda without the steradian (frei/twostream.py:46-67)."""
    T = float(U.value(temperature, 'K'))

    def bb(wavelength):
        lam_um = np.atleast_1d(U.value(wavelength, 'um'))
        c1, c2, _, _, _ = _spectral_host(lam_um, 2.4 * U.m_p, T, 1.0)
        return U.wrap(c1 / np.expm1(c2 / T), 'flux')
    return bb


def B_star(T_star, lam):
    """Blackbody spectrum of the star (frei/core.py:58-62)."""
    return BB(T_star)(lam)


def F_TOA(lam, T_star=5800, f=2 / 3, a_rstar=0.03 * U.au / U.R_sun):
    """Stellar flux at the top of the atmosphere (frei/core.py:48-55)."""
    lam_um = np.atleast_1d(U.value(lam, 'um'))
    out = _spectral_host(lam_um, 2.4 * U.m_p, float(U.value(T_star, 'K')), float(a_rstar), f)
    return U.wrap(out[4], 'flux')


class Planet(object):
    """Container for planetary system information (frei/core.py:65-106)."""

    def __init__(self, a_rstar, m_bar, g, T_star, alpha):
        self.a_rstar = a_rstar
        self.m_bar = m_bar
        self.g = g
        self.T_star = T_star
        self.alpha = alpha

    @classmethod
    def from_hot_jupiter(cls):
        """M = M_J, R = R_J, m_bar = 2.4 m_p, T_star = 5800 K, a = 0.03 au (frei/core.py:92-106)."""
        g_cgs = U.GM_jup / U.R_jup ** 2
        return cls(a_rstar=float(0.03 * U.au / U.R_sun),
                   m_bar=U.wrap(2.4 * U.m_p, 'g'),
                   g=U.wrap(g_cgs / 100.0, 'm/s2'),
                   T_star=U.wrap(5800.0, 'K'),
                   alpha=1)


class Spectrum(object):
    """Fallback for ``specutils.Spectrum1D`` (frei/core.py:335-337): flux + spectral axis."""

    def __init__(self, flux, spectral_axis):
        self.flux = flux
        self.spectral_axis = spectral_axis

    @property
    def wavelength(self):
        return self.spectral_axis


def _make_spectrum(flux, lam):
    specutils = U.optional_module('specutils')
    if specutils is not None:                               # pragma: no cover
        return specutils.Spectrum1D(flux=flux, spectral_axis=lam)
    return Spectrum(flux, lam)


def converged_layers(temp_hists, dT, n_zero_crossings, convergence_dT):
    """
    Per-layer convergence flags of the outer loop (frei/core.py:306-311): more
    than ``n_zero_crossings`` sign flips of the temperature history differences,
    or the last absorb step below ``convergence_dT``.
    """
    temp_hist = np.hstack(temp_hists)
    temp_hist = temp_hist.T[temp_hist[0] != 0].T
    diffs = np.diff(temp_hist.T, axis=0)
    flips = np.count_nonzero(np.sign(diffs[1:]) != np.sign(diffs[:-1]), axis=0)
    return (flips > n_zero_crossings) | (np.abs(dT) < convergence_dT), temp_hist


TRACE = None      # set to a list to collect (label, time.perf_counter()) marks of emission_spectrum (scripts/e2e_phases.py)


def _mark(label):
    if TRACE is not None:
        import time
        TRACE.append((label, time.perf_counter()))


class _PinnedOutputs:
    """
    Page-locked host buffers for the large results (spectrum, dtaus).  Allocating and pinning
    ~n_layers x n_lambda doubles costs more than a whole solve, so a buffer is reused once the
    arrays handed out from it have been garbage-collected by the caller.
    """

    def __init__(self):
        self._free = {}
        self._out = []

    def get(self, shape):
        import torch
        self._out = [(ref, buf) for ref, buf in self._out if ref() is not None or
                     self._free.setdefault(tuple(buf.shape), []).append(buf)]
        pool = self._free.get(tuple(shape), [])
        return pool.pop() if pool else torch.empty(shape, dtype=torch.float64).pin_memory()

    def hand_out(self, buf):
        import weakref
        arr = buf.numpy()
        self._out.append((weakref.ref(arr), buf))
        return arr


class Grid(object):
    """Grid over temperatures, pressures and wavelengths (frei/core.py:109-338)."""

    def __init__(self, planet, lam=None, pressures=None, init_temperatures=None,
                 lam_min=0.5, lam_max=10, n_wl_bins=500,
                 P_toa=1e-6, P_boa=200, n_layers=30,
                 T_ref=2300, P_ref=0.1, alpha=0.1):
        self.planet = planet
        if lam is None:
            self.lam, self.wl_bins, self.R = wavelength_grid(
                min_micron=float(U.value(lam_min, 'um')),
                max_micron=float(U.value(lam_max, 'um')), n_bins=n_wl_bins)
        else:
            self.lam, self.wl_bins, self.R = wavelength_grid(lam=lam)
        if pressures is None:
            self.pressures = pressure_grid(
                n_layers=n_layers, P_toa=np.log10(float(U.value(P_toa, 'bar'))),
                P_boa=np.log10(float(U.value(P_boa, 'bar'))))
        else:
            self.pressures = pressures
        if init_temperatures is None:
            self.init_temperatures = temperature_grid(self.pressures, T_ref, P_ref, alpha)
        else:
            self.init_temperatures = init_temperatures
        self.opacities = None
        self._table = None
        self.table_dtype = FREI_F64
        self.flux_dtype = FREI_F64          # FREI_F32: fp32 arithmetic mode (1e-4 contract)
        self._outputs = _PinnedOutputs()

    def __repr__(self):
        T = U.value(self.init_temperatures, 'K')
        P = U.value(self.pressures, 'bar')
        lam = U.value(self.lam, 'um')
        return (f"<Grid in T=[{T[0]:.0f}...{T[-1]:.0f}] K, p=[{P[0]:.2g}...{P[-1]:.2g}] bar, "
                f"lam=[{lam[0]}...{lam[-1]}] um>")

    def load_opacities(self, species=None, path=None, opacities=None, client=None,
                       force_reload=False, groupies=False):
        """
        Attach opacity tables (frei/core.py:198-231).  Pre-computed tables are
        passed with ``opacities=``; they are uploaded to the GPU on first use.
        """
        if (self.opacities is None and opacities is None) or force_reload:
            from .opacity import binned_opacity
            self.opacities = binned_opacity(self.init_temperatures, self.pressures,
                                            self.wl_bins, self.lam, species=species,
                                            groupies=groupies, path=path)
        else:
            self.opacities = opacities
        self._table = None
        return self.opacities

    # -- device residency -----------------------------------------------------
    def attach_device_table(self, table, species=None):
        """
        Use tables that already live in HBM (a :class:`~frei_b200.engine.DeviceTable`
        holding this rank's wavelength slice) instead of uploading host arrays.
        """
        self.opacities = {k: None for k in (species or table.species)}
        self._table = table
        self._table_key = 'attached'
        return self.opacities

    def device_table(self, group=None):
        if getattr(self, '_table_key', None) == 'attached':
            return self._table
        lam_range = None
        if group is not None:
            import torch.distributed as dist
            n = U.value(self.lam, 'um').shape[0]
            lam_range = shard_range(n, dist.get_rank(group), dist.get_world_size(group))
        if self._table is None or self._table_key != (lam_range, self.table_dtype):
            self._table = DeviceTable(self.opacities, dtype=self.table_dtype, lam_range=lam_range)
            self._table_key = (lam_range, self.table_dtype)
        return self._table

    def _mmr(self, T, P, m_bar_g):
        species = list(self.opacities.keys())
        d = chemistry(T, P, species, m_bar=m_bar_g)
        return np.stack([np.broadcast_to(d[s], T.shape) for s in species], axis=-1)

    def make_engine(self, group=None, want_dtaus=False):
        pl = self.planet
        T0 = U.value(self.init_temperatures, 'K')
        P = U.value(self.pressures, 'bar')
        m_bar_g = float(U.value(pl.m_bar, 'g'))
        return Engine(self.device_table(group), U.value(self.lam, 'um'), P, T0,
                      self._mmr(T0, P, m_bar_g), g=U.gravity_cgs(pl.g), m_bar=m_bar_g,
                      alpha=pl.alpha, T_star=float(U.value(pl.T_star, 'K')),
                      a_rstar=float(pl.a_rstar), group=group, want_dtaus=want_dtaus,
                      flux_dtype=self.flux_dtype, collective=getattr(self, 'collective', 'auto'))

    def emission_spectrum(self, n_timesteps=1, n_zero_crossings=2, convergence_dT=3,
                          group=None, dynamic_chemistry=None, gather='all'):
        """
        Iterate emit/absorb sweeps towards radiative equilibrium and return
        ``(spectrum, final_temps, temperature_history, dtaus)`` exactly as
        frei/core.py:233-338.  ``group`` shards the wavelength axis over a
        torch.distributed process group.  ``gather='all'`` (default): every rank
        returns the full result, as a single-process call of the reference would;
        ``gather='local'``: every rank returns only its own wavelength slice
        ``self.lam_range`` of the spectrum and of ``dtaus`` (the temperatures are
        replicated anyway) — the results of an N-GPU job then leave the GPUs over
        N PCIe links in parallel instead of N times over each of them.
        ``dynamic_chemistry``: recompute mixing ratios from the current T before
        every sweep (default: only when pyfastchem is installed; the mock's
        ratios do not depend on T).
        """
        import torch
        if self.opacities is None:
            raise ValueError("Must load opacities before computing emission spectrum.")
        if gather not in ('all', 'local'):
            raise ValueError("gather must be 'all' or 'local'")
        conv_dT = float(U.value(convergence_dT, 'K'))
        if dynamic_chemistry is None:
            dynamic_chemistry = U.optional_module('pyfastchem') is not None
        _mark('enter')
        eng = self._solver_engine(group)
        _mark('engine reset')
        P = U.value(self.pressures, 'bar')
        m_bar_g = float(U.value(self.planet.m_bar, 'g'))
        L = eng.L
        temp_hists = []
        tracked = None
        self.n_iterations = 0
        if not dynamic_chemistry and n_timesteps > 0:
            # Static mixing ratios: the convergence rule of frei/core.py:301-318 runs on the device
            # (Engine.enable_batch_convergence, the tracker of the batch mode with B = 1): once
            # it fires, every later kernel of this atmosphere exits at once, so iterations can be
            # queued back to back and the host looks at the flag only every `check_every`
            # iterations instead of synchronising after each one.  The temperature history is
            # written by the update kernel into a device buffer, one [2][L] slot per iteration.
            # The host never blocks on a flag: every `check_every` iterations it queues a copy of
            # the flag into pinned memory behind an event and only looks at copies whose event has
            # completed, while it keeps the device fed (at most `max_ahead` iterations beyond the
            # oldest unread flag; iterations queued after the rule has fired cost a launch each).
            check_every, max_ahead = 4, 16
            eng.enable_batch_convergence(n_zero_crossings, conv_dT)
            hist = eng.history_buffer(n_timesteps)             # [n_timesteps][2][B][L], reused across solves
            flags = eng.flag_ring(max_ahead // check_every + 2)
            main = torch.cuda.current_stream(eng.device)
            side = eng.side_stream()                           # flag copies stay off the sweeps' stream
            pending, it, stopped = collections.deque(), 0, False
            while it < n_timesteps and not stopped:
                eng.sweep(FREI_EMIT, T_hist=hist[it, 0])
                eng.sweep(FREI_ABSORB, T_hist=hist[it, 1])
                it += 1
                if it % check_every == 0 and it < n_timesteps:
                    slot = (it // check_every) % flags.shape[0]
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        eng.active.record_stream(side)
                        flags[slot].copy_(eng.active[:1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                    pending.append((ev, slot, it))
                if pending and it - pending[0][2] >= max_ahead:
                    pending[0][0].synchronize()
                while pending and pending[0][0].query():
                    _, slot, _ = pending.popleft()
                    if int(flags[slot, 0]) == 0:
                        stopped = True
            _mark('iterations queued')
            # No synchronisation here: a converged atmosphere ignores the sweeps queued behind its
            # last one, so the final emit and the copies of the results follow at once and the host
            # waits a single time, for everything (below).
            tracked = (it, eng.active, eng.iterations_done, hist)
            eng.disable_batch_convergence()
        for it in range(n_timesteps if dynamic_chemistry else 0):
            if it > 0:
                eng.set_mmr(self._mmr(eng.get_T()[0], P, m_bar_g))
            eng.sweep(FREI_EMIT, T_hist=eng.hist[0])
            eng.set_mmr(self._mmr(eng.get_T()[0], P, m_bar_g))
            eng.sweep(FREI_ABSORB, T_hist=eng.hist[1])
            T_emit, T_absorb, dT = eng.read_history()
            temp_hists.append(np.stack([T_emit[0], T_absorb[0]], axis=1))
            dT = dT[0]
            self.n_iterations += 1
            conv, _ = converged_layers(temp_hists, dT, n_zero_crossings, conv_dT)
            if np.all(conv):
                break
        if dynamic_chemistry and n_timesteps > 0:
            eng.set_mmr(self._mmr(eng.get_T()[0], P, m_bar_g))
        # final emit: alpha is not forwarded -> default 1 (frei/core.py:323-333)
        eng.sweep(FREI_EMIT, alpha_override=1.0, with_dtaus=True)
        _mark('final emit queued')
        spec_local = eng.F_up[0, L - 1]
        dtaus_local = eng.dtaus[0]
        lam_out = self.lam
        self.lam_range = (0, eng.n_lam_global)
        if group is not None and gather == 'all':
            from .sharding import gather_lambda
            both = torch.cat([spec_local[None, :], dtaus_local], dim=0)     # [L + 1][n_local]
            full = gather_lambda(both, eng.n_lam_global, group, to_numpy=False)
        else:
            full = None
            if group is not None:                              # this rank's slice only
                self.lam_range = (eng.lo, eng.hi)
                lam_out = U.wrap(U.value(self.lam, 'um')[self.lam_range[0]:self.lam_range[1]], 'um')
        n_out = self.lam_range[1] - self.lam_range[0]
        out = self._outputs.get((L + 1, n_out))                # pinned: [0] spectrum, [1:] dtaus
        if full is not None:
            out.copy_(full, non_blocking=True)
        else:
            out[0].copy_(spec_local, non_blocking=True)
            out[1:].copy_(dtaus_local, non_blocking=True)
        small = eng.pinned_results(n_timesteps)                # T, history, flag, iteration count
        small['T'].copy_(eng.T[0], non_blocking=True)
        if tracked is not None:
            n_q, active_dev, iters_dev, hist = tracked
            if n_q:
                small['hist'][:n_q].copy_(hist[:n_q, :, 0, :], non_blocking=True)
            small['active'].copy_(active_dev[:1], non_blocking=True)
            small['iters'].copy_(iters_dev[:1], non_blocking=True)
        _mark('copies queued')
        torch.cuda.current_stream(eng.device).synchronize()    # the one wait of a solve
        if tracked is not None:
            eng.side_stream().synchronize()                    # flag polls still in flight (they only depend on
                                                               # iterations that are long done): the ring is reused
        _mark('results on host')
        eng.check_errors()
        final_temps = small['T'].numpy().copy()
        if tracked is not None:
            self.n_iterations = n_q if int(small['active'][0]) else int(small['iters'][0])
            h = small['hist'][:self.n_iterations].numpy()                             # [n][2][L]
            temp_hists = [h[k].T.copy() for k in range(self.n_iterations)]
        temp_hist = np.hstack(temp_hists) if temp_hists else np.zeros((L, 0))
        if temp_hists:
            temp_hist = temp_hist.T[temp_hist[0] != 0].T
        arr = self._outputs.hand_out(out)
        spec, dtaus = arr[0], arr[1:]
        self.engine = eng
        res = (_make_spectrum(U.wrap(spec, 'flux'), lam_out), U.wrap(final_temps, 'K'),
               U.wrap(temp_hist, 'K'), dtaus)
        _mark('return')
        return res

    def _solver_engine(self, group):
        """Device state of this Grid, built once and reset for every solve."""
        pl = self.planet
        T0 = U.value(self.init_temperatures, 'K')
        P = U.value(self.pressures, 'bar')
        m_bar_g = float(U.value(pl.m_bar, 'g'))
        table = self.device_table(group)
        key = (id(table), id(group), T0.shape, U.value(self.lam, 'um').shape,
               U.gravity_cgs(pl.g), m_bar_g, float(pl.alpha), float(U.value(pl.T_star, 'K')),
               float(pl.a_rstar), P.tobytes(), self.flux_dtype, getattr(self, 'collective', 'auto'))
        if getattr(self, '_eng_key', None) != key:
            self._eng = self.make_engine(group=group, want_dtaus=True)
            self._eng_key = key
        else:
            self._eng.reset(T0, self._mmr(T0, P, m_bar_g))
        return self._eng

    def diagnostics(self, contribution_function=False, pressure_milne=False, group=None):
        """
        T_eff (frei/core.py:386-439) and optionally the contribution function
        (frei/plot.py:63-79) of the last ``emission_spectrum`` solve, computed from the
        spectrum and ``dtaus`` still resident in HBM — no host copies of the large arrays.
        With a wavelength-sharded ``group`` the three sums are added over the ranks and the
        per-wavelength outputs cover this rank's slice ``self.lam_range``.
        Returns a dict with ``T_eff``, ``T_milne``, ``T_planck`` [K] and, on request,
        ``pressure_milne`` [n_lambda] and ``contribution_function`` [n_layers][n_lambda]
        (numpy arrays).
        """
        eng = getattr(self, 'engine', None)
        if eng is None or eng.dtaus is None:
            raise ValueError("Must run emission_spectrum before computing its diagnostics.")
        L = eng.L
        sums, pm, cf = _device_diagnostics(
            eng.lam_dev[eng.lo:eng.hi], eng.w, eng.P[0], eng.T[0], eng.F_up[0, L - 1], eng.dtaus[0],
            want_pressure=pressure_milne, want_cf=contribution_function)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        P, T = eng.P[0].cpu().numpy(), eng.T[0].cpu().numpy()
        t_milne, t_planck = _teff_from_sums(sums.cpu().numpy(), P, T)
        out = {'T_eff': U.wrap(np.mean([t_milne, t_planck]), 'K'), 'T_milne': U.wrap(t_milne, 'K'),
               'T_planck': U.wrap(t_planck, 'K')}
        if pm is not None:
            out['pressure_milne'] = pm.cpu().numpy()
        if cf is not None:
            out['contribution_function'] = cf.cpu().numpy()
        return out

    def emission_dashboard(self, *args, **kwargs):
        raise NotImplementedError('plotting is outside the scope of frei_b200 (frei/plot.py)')


# -- T_eff diagnostics (frei/core.py:386-439) and contribution function (frei/plot.py:63-79) ------
# on the device (frei_b200_diagnostics): the reference loops over the wavelengths in Python, one
# np.interp call each — about a second for 200k bins; here the spectrum and dtaus of a solve
# stay in HBM and three sums come back.
def _trapz_weights_cm(lam_um):
    lam_cm = np.asarray(lam_um, dtype=np.float64) * 1e-4
    w = np.zeros_like(lam_cm)
    if lam_cm.shape[0] > 1:
        w[1:-1] = 0.5 * (lam_cm[2:] - lam_cm[:-2])
        w[0] = 0.5 * (lam_cm[1] - lam_cm[0])
        w[-1] = 0.5 * (lam_cm[-1] - lam_cm[-2])
    return w


def _device_diagnostics(lam_um, w_cm, P_bar, T, spec, dtaus, want_pressure=False, want_cf=False):
    """
    Run ``frei_b200_diagnostics`` on arrays that are numpy (uploaded) or torch device tensors
    (used in place).  Returns ``(sums[3] device tensor, pressure_milne or None, cf or None)``.
    """
    import torch
    lib = _cabi.load()
    _cabi.require_cuda()
    dev = None
    for x in (dtaus, spec, lam_um):
        if isinstance(x, torch.Tensor) and x.is_cuda:
            dev = x.device
    if dev is None:
        dev = torch.device('cuda', torch.cuda.current_device())

    def dbl(x):
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
        return x.to(device=dev, dtype=torch.float64).contiguous()

    d_dtaus, d_spec, d_lam, d_w, d_P, d_T = (dbl(x) for x in (dtaus, spec, lam_um, w_cm, P_bar, T))
    L, n = d_dtaus.shape
    if d_spec.shape != (n,) or d_lam.shape != (n,) or d_w.shape != (n,) or d_P.shape != (L,) or d_T.shape != (L,):
        raise ValueError('inconsistent shapes: dtaus [L][n_lam], spectrum/lam [n_lam], pressures/temperatures [L]')
    scratch = torch.empty(max(1, lib.frei_b200_diagnostics_scratch_bytes(n) // 8), dtype=torch.float64, device=dev)
    sums = torch.empty(3, dtype=torch.float64, device=dev)
    pm = torch.empty(n, dtype=torch.float64, device=dev) if want_pressure else None
    cf = torch.empty((L, n), dtype=torch.float64, device=dev) if want_cf else None
    with torch.cuda.device(dev):
        _cabi.check(lib.frei_b200_diagnostics(
            d_dtaus.data_ptr(), d_spec.data_ptr(), d_lam.data_ptr(), d_w.data_ptr(), d_P.data_ptr(),
            d_T.data_ptr(), L, n, _cabi.ptr(pm), _cabi.ptr(cf), scratch.data_ptr(), sums.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream))
    return sums, pm, cf


def _teff_from_sums(sums, P_bar, T):
    """(T_milne, T_planck) from the three device sums (frei/core.py:397-403, 413-414)."""
    s = [float(x) for x in sums]
    p_avg = s[0] / s[1]
    P_bar, T = np.asarray(P_bar, dtype=np.float64), np.asarray(T, dtype=np.float64)
    return float(np.interp(p_avg, P_bar[::-1], T[::-1])), float((s[2] / U.sigma_sb) ** (1 / 4))


def _grid_arrays(grid, spec, final_temps):
    lam_um = U.value(grid.lam, 'um')
    lo, hi = getattr(grid, 'lam_range', None) or (0, lam_um.shape[0])
    flux = U.value(spec.flux, 'flux')
    w = _trapz_weights_cm(lam_um)
    if flux.shape[0] != lam_um.shape[0]:          # a rank's slice (gather='local')
        lam_um, w = lam_um[lo:hi], w[lo:hi]
    return lam_um, w, U.value(grid.pressures, 'bar'), U.value(final_temps, 'K'), flux


def effective_temperature_milne(grid, spec, dtaus, final_temps):
    """Photosphere temperature from Milne's tau ~ 2/3 (frei/core.py:386-405)."""
    lam_um, w, P, T, flux = _grid_arrays(grid, spec, final_temps)
    sums, _, _ = _device_diagnostics(lam_um, w, P, T, flux, dtaus)
    return U.wrap(_teff_from_sums(sums.cpu().numpy(), P, T)[0], 'K')


def effective_temperature_planck(grid, spec):
    """Invert the Stefan-Boltzmann law for the emitted bolometric flux (frei/core.py:408-414)."""
    lam_cm = U.value(grid.lam, 'um') * 1e-4
    flux = U.value(spec.flux, 'flux')
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    return U.wrap((trapz(flux, lam_cm) / U.sigma_sb) ** (1 / 4), 'K')


def effective_temperature(grid, spec, dtaus, final_temps):
    """Mean of the Milne and Stefan-Boltzmann estimates (frei/core.py:417-439)."""
    lam_um, w, P, T, flux = _grid_arrays(grid, spec, final_temps)
    sums, _, _ = _device_diagnostics(lam_um, w, P, T, flux, dtaus)
    a, b = _teff_from_sums(sums.cpu().numpy(), P, T)
    return U.wrap(np.mean([a, b]), 'K')


def contribution_function(grid, dtaus, final_temps):
    """
    Normalised contribution function [n_layers][n_lambda] in level order, as the dashboard
    shows it (``cf[::-1]`` of frei/plot.py:63-83).
    """
    lam_um = U.value(grid.lam, 'um')
    P, T = U.value(grid.pressures, 'bar'), U.value(final_temps, 'K')
    n = np.asarray(dtaus).shape[1]
    lo, hi = (0, n) if n == lam_um.shape[0] else grid.lam_range
    zeros = np.zeros(n)
    _, _, cf = _device_diagnostics(lam_um[lo:hi], zeros, P, T, zeros, dtaus, want_cf=True)
    return cf.cpu().numpy()
