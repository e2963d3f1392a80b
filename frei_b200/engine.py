"""
Device-resident state of the radiative-equilibrium hot path and the thin layer
that drives the C ABI (``include/frei_b200.h``).  PyTorch is used only for
device buffers, streams and ``torch.distributed``; all arithmetic on the path
runs in the CUDA kernels of ``csrc/frei_b200.cu``.

Layout in HBM (one GPU):
  table   [S][N_P][N_T][n_lam_local]   fp32 or fp64, wavelength contiguous
  F_up, F_down (, dtaus)  [B][L][n_lam_local]   fp64
  c1, c2, sigma, w, f_toa [n_lam_local] fp64
  T, P [B][L]; mmr [B][L][S]; g, m_bar, alpha [B]        fp64
Wavelength sharding: every rank holds a contiguous slice of the wavelength axis
of every array above; T/P/mmr are replicated and the [B][L][4] wavelength
integrals are summed across ranks once per sweep.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import FREI_EMIT, FREI_ABSORB, FREI_F32, FREI_F64
from .sharding import shard_range, allreduce_sums

__all__ = ['DeviceTable', 'Engine', 'FREI_EMIT', 'FREI_ABSORB', 'shard_range']


def _torch():
    import torch
    return torch


def normalise_table(tab):
    """
    Bring one species' table to (P_axis ascending [bar], T_axis ascending [K],
    values[N_P, N_T, n_lam], has_T).  Accepts the oracle-style dict(P=, T=,
    values=) or an xarray.DataArray-like object with ``pressure``,
    ``temperature`` coords and ``dims`` (frei/opacity.py:331-339, 467-477).
    """
    if isinstance(tab, dict):
        Pax = np.asarray(tab['P'], dtype=np.float64)
        Tax = np.asarray(tab['T'], dtype=np.float64)
        vals = np.asarray(tab['values'])
    else:
        Pax = np.asarray(getattr(tab.pressure, 'values', tab.pressure), dtype=np.float64)
        Tax = np.asarray(getattr(tab.temperature, 'values', tab.temperature), dtype=np.float64)
        Tax = np.asarray(getattr(Tax, 'value', Tax), dtype=np.float64)
        vals = np.asarray(getattr(tab, 'values'))
        dims = tuple(getattr(tab, 'dims', ('pressure', 'temperature', 'wavelength')))
        order = [dims.index(d) for d in ('pressure', 'temperature', 'wavelength')]
        vals = np.transpose(vals, order)
    ip = np.argsort(Pax, kind='stable')
    it = np.argsort(Tax, kind='stable')
    Pax, Tax = Pax[ip], Tax[it]
    vals = vals[ip][:, it]
    has_T = len(np.unique(Tax)) > 1                    # frei/opacity.py:256
    if not has_T:
        Tax = np.array([Tax[0], Tax[0] + 1.0])
        vals = np.repeat(vals[:, :1], 2, axis=1)
    return Pax, Tax, vals, has_T


class DeviceTable:
    """Opacity tables of all species as one dense device block (see ``frei_table``)."""

    def __init__(self, opacities, device=None, dtype=FREI_F64, lam_range=None):
        torch = _torch()
        _cabi.require_cuda()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        items = list(opacities.items()) if isinstance(opacities, dict) else \
            [(str(i), t) for i, t in enumerate(opacities)]
        self.species = [k for k, _ in items]
        norm = [normalise_table(t) for _, t in items]
        N_P = {n[0].shape[0] for n in norm}
        N_T = {n[1].shape[0] for n in norm if n[3]}
        if len(N_P) != 1 or len(N_T) > 1:
            raise ValueError('all species tables must share the (pressure, temperature) grid sizes')
        self.N_P = N_P.pop()
        self.N_T = N_T.pop() if N_T else 2
        self.S = len(norm)
        n_lam_global = norm[0][2].shape[2]
        lo, hi = (0, n_lam_global) if lam_range is None else lam_range
        self.lam_range = (lo, hi)
        self.n_lam = hi - lo
        self.dtype = dtype
        tdt = torch.float32 if dtype == FREI_F32 else torch.float64
        ndt = np.float32 if dtype == FREI_F32 else np.float64
        self.values = torch.empty((self.S, self.N_P, self.N_T, self.n_lam), dtype=tdt,
                                  device=self.device)
        axP = np.empty((self.S, self.N_P))
        axT = np.empty((self.S, self.N_T))
        hasT = np.empty(self.S, dtype=np.int32)
        for s, (Pax, Tax, vals, has_T) in enumerate(norm):
            if not has_T and self.N_T > 2:
                Tax = Tax[0] + np.arange(self.N_T, dtype=np.float64)
                vals = np.repeat(vals[:, :1], self.N_T, axis=1)
            axP[s], axT[s], hasT[s] = Pax, Tax, int(has_T)
            chunk = np.ascontiguousarray(vals[:, :, lo:hi], dtype=ndt)
            self.values[s].copy_(torch.from_numpy(chunk))
        self.axis_P = torch.from_numpy(axP).to(self.device)
        self.axis_T = torch.from_numpy(axT).to(self.device)
        self.has_T = torch.from_numpy(hasT).to(self.device)

    @classmethod
    def from_device_values(cls, values, axis_P, axis_T, has_T=None, species=None):
        """Wrap an existing device tensor [S][N_P][N_T][n_lam] (no copy)."""
        torch = _torch()
        self = cls.__new__(cls)
        self.device = values.device
        self.S, self.N_P, self.N_T, self.n_lam = values.shape
        self.lam_range = (0, self.n_lam)
        self.dtype = FREI_F32 if values.dtype == torch.float32 else FREI_F64
        self.values = values.contiguous()
        axP = np.broadcast_to(np.asarray(axis_P, dtype=np.float64), (self.S, self.N_P))
        axT = np.broadcast_to(np.asarray(axis_T, dtype=np.float64), (self.S, self.N_T))
        self.axis_P = torch.from_numpy(np.ascontiguousarray(axP)).to(self.device)
        self.axis_T = torch.from_numpy(np.ascontiguousarray(axT)).to(self.device)
        hT = np.ones(self.S, dtype=np.int32) if has_T is None else np.asarray(has_T, dtype=np.int32)
        self.has_T = torch.from_numpy(hT).to(self.device)
        self.species = species or [str(i) for i in range(self.S)]
        return self

    def struct(self):
        return _cabi.frei_table(self.values.data_ptr(), self.axis_P.data_ptr(),
                                self.axis_T.data_ptr(), self.has_T.data_ptr(),
                                self.S, self.N_P, self.N_T, self.dtype, self.n_lam)


class Engine:
    """
    B atmospheres x L levels x n_lam wavelengths resident on one GPU.

    Parameters are plain numbers/arrays in the units of ``include/frei_b200.h``
    (T [K], P [bar], g [cm s^-2], m_bar [g], wavelength [micron]).
    ``lam_um`` is the GLOBAL grid; ``table`` holds the slice ``table.lam_range``.
    ``group``: a torch.distributed process group over which the wavelength
    axis is sharded (None = single device / independent batches).
    """

    def __init__(self, table, lam_um, pressures_bar, temperatures, mmr, g, m_bar, alpha=1.0,
                 T_star=5800.0, a_rstar=1.0, f_toa=None, ftoa_scale=None, group=None,
                 flux_dtype=FREI_F64, want_dtaus=False, collective='auto'):
        torch = _torch()
        self.lib = _cabi.load()
        _cabi.require_cuda()
        self.table = table
        self.device = dev = table.device
        self.group = group
        f64 = torch.float64
        P = np.atleast_2d(np.asarray(pressures_bar, dtype=np.float64))
        T = np.atleast_2d(np.asarray(temperatures, dtype=np.float64))
        self.B, self.L = P.shape
        if self.L < 3:
            raise ValueError('need at least 3 levels')
        self.S = table.S
        lam_um = np.ascontiguousarray(lam_um, dtype=np.float64)
        self.n_lam_global = lam_um.shape[0]
        self.lo, self.hi = table.lam_range
        self.n_lam = table.n_lam
        B, L, S, n = self.B, self.L, self.S, self.n_lam

        def dvec(x, shape):
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), shape))
            return torch.from_numpy(a.copy()).to(dev)

        self.P = dvec(P, (B, L))
        self.T = dvec(T, (B, L))
        self.mmr = dvec(mmr, (B, L, S))
        self.g = dvec(g, (B,))
        m_bar_b = np.broadcast_to(np.asarray(m_bar, dtype=np.float64), (B,))
        self.m_bar = dvec(m_bar_b, (B,))
        self.alpha = dvec(alpha, (B,))
        self.sigma_scale = None
        if not np.all(m_bar_b == m_bar_b[0]):
            self.sigma_scale = dvec(m_bar_b[0] / m_bar_b, (B,))
        self.ftoa_scale = None if ftoa_scale is None else dvec(ftoa_scale, (B,))

        # per-wavelength constants (K: spectral_setup)
        self.lam_dev = torch.from_numpy(lam_um).to(dev)
        self.c1, self.c2, self.sigma, self.w, self.f_toa = (
            torch.empty(n, dtype=f64, device=dev) for _ in range(5))
        _cabi.check(self.lib.frei_b200_spectral_setup(
            self.lam_dev.data_ptr(), self.n_lam_global, self.lo, n, float(m_bar_b[0]),
            float(T_star), float(a_rstar), 2.0 / 3.0, self.c1.data_ptr(), self.c2.data_ptr(),
            self.sigma.data_ptr(), self.w.data_ptr(), self.f_toa.data_ptr(), self._stream()))
        if f_toa is not None:       # caller-supplied F_TOA (emit()/absorb() signature)
            f_toa = np.ascontiguousarray(f_toa, dtype=np.float64)
            self.f_toa.copy_(torch.from_numpy(f_toa[self.lo:self.hi]))

        # flux state
        # FREI_F64: fp64 arithmetic (parity 1e-6 contract); FREI_F32: fp32 state and arithmetic
        # with fp64 wavelength integrals (1e-4 contract)
        self.flux_dtype = flux_dtype
        self._fdt = f64 if flux_dtype == FREI_F64 else torch.float32
        self.F_up = torch.zeros((B, L, n), dtype=self._fdt, device=dev)
        self.F_down = torch.zeros((B, L, n), dtype=self._fdt, device=dev)
        self.dtaus = torch.empty((B, L, n), dtype=self._fdt, device=dev) if want_dtaus else None

        # workspace
        sizes = [C.c_int64() for _ in range(4)]
        _cabi.check(self.lib.frei_b200_workspace_bytes(B, L, S, n, *[C.byref(s) for s in sizes]))
        self._lp = torch.empty(sizes[0].value, dtype=torch.uint8, device=dev)
        self._partials = torch.zeros((sizes[1].value + 7) // 8, dtype=f64, device=dev)   # incl. ticket counters
        self.sums = torch.zeros((B, L, 4), dtype=f64, device=dev)
        # [0] = T after the last emit, [1] = T after the last absorb, [2] = dT of the last sweep:
        # one small D2H copy per iteration brings back everything the host convergence test needs
        self.hist = torch.zeros((3, B, L), dtype=f64, device=dev)
        self.dT = self.hist[2]
        self._hist_host = None
        self.launches = 0
        self._p2p = None
        self._graph = None
        if group is not None and collective in ('auto', 'p2p'):
            self._setup_p2p(required=(collective == 'p2p'))
        self._records_stale = True     # level records must be rebuilt before the next sweep
        self.sweep_events = None        # list of (start, end) CUDA events around the sweep kernel
        self._build_structs()

    # -- fused cross-GPU sum over peer memory --------------------------------
    def _setup_p2p(self, required=False):
        """
        Exchange buffers in symmetric (peer-mapped) memory for the one-launch
        reduce + all-reduce + update (``frei_b200_post_p2p``).  All ranks must end up with the
        same collective: a rank whose rendezvous failed while its peers spin on words that never
        arrive would hang the job, so the outcome is agreed on with an all-reduce (MIN) and every
        rank falls back to the NCCL all-reduce between two launches together — with a warning.
        """
        torch = _torch()
        import warnings
        import torch.distributed as dist
        p2p, why = None, ''
        try:
            import torch.distributed._symmetric_memory as symm
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            words = 2 * world * self.B * self.L * 4 * 2        # [2][world][B][L*4][2] 8-byte words
            buf = symm.empty(words, dtype=torch.int64, device=self.device)
            buf.zero_()
            hb = symm.rendezvous(buf, self.group)
            torch.cuda.synchronize(self.device)
            p2p = dict(buf=buf, hb=hb, rank=rank, world=world, epoch=0,
                       bufs=torch.tensor(list(hb.buffer_ptrs), dtype=torch.int64, device=self.device),
                       err=torch.zeros(1, dtype=torch.int32, device=self.device))
        except Exception as exc:                               # no symmetric memory on this system
            why = repr(exc)
        ok = torch.tensor([1 if p2p is not None else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)       # also the barrier after zero_()
        if int(ok.item()) == 1:
            self._p2p = p2p
            return
        self._p2p = None
        msg = ('frei_b200: peer-memory exchange unavailable on at least one rank'
               + (f' (this rank: {why})' if why else '') + '; all ranks use the NCCL all-reduce')
        if required:
            raise _cabi.FreiError(msg)
        warnings.warn(msg)

    def check_errors(self):
        """
        Raise if the fused cross-GPU exchange timed out on this rank (a peer did not deliver its
        integrals): the sums, and every temperature since, are invalid.  Synchronises the stream;
        called wherever the host already waits for the device.
        """
        if self._p2p is not None and int(self._p2p['err'].item()) != 0:
            raise _cabi.FreiError('frei_b200: peer-memory all-reduce timed out waiting for a rank; '
                                  'temperatures on this rank are invalid')

    # -- plumbing -----------------------------------------------------------
    def _stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    def _build_structs(self):
        self._graph = None          # a captured iteration holds the old structs' pointers
        n = self.n_lam
        self._tab = self.table.struct()
        self._spec = _cabi.frei_spectral(self.c1.data_ptr(), self.c2.data_ptr(),
                                         self.sigma.data_ptr(), self.w.data_ptr(),
                                         self.f_toa.data_ptr(), n)
        trk = getattr(self, '_tracker', None)
        self._atm = _cabi.frei_atmosphere(
            self.T.data_ptr(), self.P.data_ptr(), self.mmr.data_ptr(), self.g.data_ptr(),
            self.m_bar.data_ptr(), self.alpha.data_ptr(), _cabi.ptr(self.sigma_scale),
            _cabi.ptr(self.ftoa_scale), self.B, self.L,
            _cabi.ptr(getattr(self, 'active', None)), C.pointer(trk) if trk is not None else None)
        self._ws = _cabi.frei_workspace(self._lp.data_ptr(), self._partials.data_ptr(),
                                        self.sums.data_ptr(), self.dT.data_ptr())

    def _flux_struct(self, with_dtaus):
        return _cabi.frei_flux(self.F_up.data_ptr(), self.F_down.data_ptr(),
                               _cabi.ptr(self.dtaus) if with_dtaus else None, self.flux_dtype)

    # -- state --------------------------------------------------------------
    def set_T(self, T):
        torch = _torch()
        self._records_stale = True
        self.T.copy_(torch.from_numpy(np.array(          # np.array: a writable, contiguous copy
            np.broadcast_to(np.asarray(T, dtype=np.float64), (self.B, self.L)))))

    def set_mmr(self, mmr):
        torch = _torch()
        self._records_stale = True
        self.mmr.copy_(torch.from_numpy(np.array(
            np.broadcast_to(np.asarray(mmr, dtype=np.float64), (self.B, self.L, self.S)))))

    def set_fluxes(self, F_up=None, F_down=None):
        """Upload the local wavelength slice of host flux arrays [B?][L][n_lam_global]."""
        torch = _torch()
        for dst, src in ((self.F_up, F_up), (self.F_down, F_down)):
            if src is not None:
                a = np.asarray(src, dtype=np.float64).reshape((self.B, self.L, -1))
                dst.copy_(torch.from_numpy(np.ascontiguousarray(a[:, :, self.lo:self.hi])).to(dst.dtype))

    def get_T(self):
        return self.T.cpu().numpy()

    def reset(self, temperatures, mmr=None):
        """Back to the initial state of a solve: zero fluxes (frei/core.py:265-266), given T."""
        self.F_up.zero_()
        self.F_down.zero_()
        self.set_T(temperatures)
        if mmr is not None:
            self.set_mmr(mmr)

    def iteration(self):
        """emit + absorb (one pass of the loop body of frei/core.py:273-299), asynchronous."""
        if self._graph is not None and not self._records_stale and self.sweep_events is None:
            self._graph.replay()
            self.launches += 4
            return
        self.sweep(FREI_EMIT, T_hist=self.hist[0])
        self.sweep(FREI_ABSORB, T_hist=self.hist[1])

    # -- batches of atmospheres: convergence decided on the device ---------------------------
    def enable_batch_convergence(self, n_zero_crossings=2, convergence_dT=3.0):
        """
        Track the convergence rule of Grid.emission_spectrum (frei/core.py:301-318) per
        atmosphere on the device; converged atmospheres are skipped by every later kernel.
        """
        torch = _torch()
        dev, B, L = self.device, self.B, self.L
        self.active = torch.ones(B, dtype=torch.uint8, device=dev)
        self._trk_T = torch.zeros((B, L), dtype=torch.float64, device=dev)
        self._trk_state = torch.zeros((B, L, 2), dtype=torch.int32, device=dev)
        self._trk_state[:, :, 0] = 2
        self._trk_ncol = torch.zeros(B, dtype=torch.int32, device=dev)
        self.iterations_done = torch.zeros(B, dtype=torch.int32, device=dev)
        self._tracker = _cabi.frei_tracker(self._trk_T.data_ptr(), self._trk_state.data_ptr(),
                                           self._trk_ncol.data_ptr(),
                                           self.iterations_done.data_ptr(), int(n_zero_crossings),
                                           float(convergence_dT))
        self._build_structs()

    def disable_batch_convergence(self):
        self.active = None
        self._tracker = None
        self._build_structs()

    def solve_batch(self, n_timesteps, n_zero_crossings=2, convergence_dT=3.0, check_every=4):
        """
        Grid.emission_spectrum for every atmosphere of the batch (frei/core.py:263-333): iterate
        emit/absorb until each atmosphere meets the convergence rule (or n_timesteps), then the
        final emit with alpha = 1.  Returns (iterations[B], T[B][L]) as host arrays; the spectra
        are F_up[:, L-1, :] on the device.  The host only polls the per-atmosphere flags every
        ``check_every`` iterations.
        """
        self.enable_batch_convergence(n_zero_crossings, convergence_dT)
        n_done = 0
        for it in range(n_timesteps):
            self.iteration()
            n_done += 1
            if (it + 1) % check_every == 0:
                done = not bool(self.active.any().item())
                self.check_errors()
                if done:
                    break
        self.check_errors()
        iters = self.iterations_done.cpu().numpy().copy()
        still = self.active.cpu().numpy().astype(bool)
        iters[still] = n_done                                  # stopped by n_timesteps
        self.disable_batch_convergence()
        self.sweep(FREI_EMIT, alpha_override=1.0)              # final emit, alpha not forwarded (:323-333)
        return iters, self.T.cpu().numpy()

    def capture_iteration(self):
        """
        Record emit + absorb (4 kernel launches on one device) into a CUDA graph so that an
        iteration is one graph launch.  Single-device engines only; the state the graph reads
        and writes (T, fluxes, records, hist) keeps its addresses for the life of the engine.
        """
        torch = _torch()
        if self.group is not None:
            return False
        if self._records_stale:
            self.layer_prep()
        # the captured launches run for real once during warm-up below; keep the state intact
        keep = [t.clone() for t in (self.T, self.F_up, self.F_down, self.hist, self._lp)]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.sweep(FREI_EMIT, T_hist=self.hist[0])      # warm-up outside capture (attribute calls)
            self.sweep(FREI_ABSORB, T_hist=self.hist[1])
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.sweep(FREI_EMIT, T_hist=self.hist[0])
            self.sweep(FREI_ABSORB, T_hist=self.hist[1])
        for dst, src in zip((self.T, self.F_up, self.F_down, self.hist, self._lp), keep):
            dst.copy_(src)
        self._graph = graph
        return True

    def history_buffer(self, n_iterations):
        """Device buffer [n][2][B][L] for the temperature history of up to n iterations (cached)."""
        torch = _torch()
        buf = getattr(self, '_hist_buf', None)
        if buf is None or buf.shape[0] < n_iterations:
            buf = torch.empty((max(n_iterations, 32), 2, self.B, self.L), dtype=torch.float64,
                              device=self.device)
            self._hist_buf = buf
        return buf

    def flag_ring(self, n):
        """Pinned host ring [n][1] of uint8 for asynchronous copies of the convergence flag."""
        torch = _torch()
        ring = getattr(self, '_flag_ring', None)
        if ring is None or ring.shape[0] < n:
            ring = torch.ones((n, 1), dtype=torch.uint8).pin_memory()
            self._flag_ring = ring
        ring.fill_(1)
        return ring

    def side_stream(self):
        """A second stream of this device for small copies that must not sit between two sweeps."""
        torch = _torch()
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def pinned_results(self, n_iterations):
        """Pinned host buffers for the small results of a solve (first atmosphere): T [L], history
        [n][2][L], the convergence flag and the iteration count (cached)."""
        torch = _torch()
        res = getattr(self, '_pinned_small', None)
        if res is None or res['hist'].shape[0] < n_iterations:
            res = dict(T=torch.empty(self.L, dtype=torch.float64).pin_memory(),
                       hist=torch.empty((max(n_iterations, 32), 2, self.L), dtype=torch.float64).pin_memory(),
                       active=torch.ones(1, dtype=torch.uint8).pin_memory(),
                       iters=torch.zeros(1, dtype=torch.int32).pin_memory())
            self._pinned_small = res
        return res

    def read_history(self):
        """(T after emit, T after absorb, dT of the absorb sweep) as host arrays [B][L]; syncs."""
        torch = _torch()
        if self._hist_host is None:
            self._hist_host = torch.empty((3, self.B, self.L), dtype=torch.float64).pin_memory()
        self._hist_host.copy_(self.hist, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.check_errors()
        h = self._hist_host.numpy()
        return h[0].copy(), h[1].copy(), h[2].copy()

    # -- kernels ------------------------------------------------------------
    def layer_prep(self, debug=False):
        torch = _torch()
        out = None
        args = [None] * 5
        if debug:
            shp = (self.B, self.L, self.S)
            out = dict(iP=torch.empty(shp, dtype=torch.int32, device=self.device),
                       iT=torch.empty(shp, dtype=torch.int32, device=self.device),
                       wP=torch.empty(shp, dtype=torch.float64, device=self.device),
                       wT=torch.empty(shp, dtype=torch.float64, device=self.device),
                       oob=torch.empty(shp, dtype=torch.uint8, device=self.device))
            args = [out[k].data_ptr() for k in ('iP', 'iT', 'wP', 'wT', 'oob')]
        _cabi.check(self.lib.frei_b200_layer_prep(C.byref(self._tab), C.byref(self._atm),
                                                  C.byref(self._ws), *args, self._stream()))
        self.launches += 1
        self._records_stale = False
        return out

    def kappa(self):
        """k[B][L][n_lam] and sigma[B][n_lam] at the current (T, P) of every level."""
        torch = _torch()
        self.layer_prep()
        k = torch.empty((self.B, self.L, self.n_lam), dtype=torch.float64, device=self.device)
        sg = torch.empty((self.B, self.n_lam), dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.frei_b200_kappa(C.byref(self._tab), C.byref(self._spec),
                                             C.byref(self._atm), C.byref(self._ws),
                                             k.data_ptr(), sg.data_ptr(), self._stream()))
        self.launches += 1
        return k, sg

    def sweep(self, direction, alpha_override=-1.0, with_dtaus=False, T_hist=None):
        """
        One emit (FREI_EMIT) or absorb (FREI_ABSORB) pass: brackets, layer sweep,
        wavelength integrals, (cross-rank sum,) temperature update.  Everything
        is enqueued on the current stream; nothing is synchronised.
        """
        if with_dtaus and self.dtaus is None:
            torch = _torch()
            self.dtaus = torch.empty((self.B, self.L, self.n_lam), dtype=self._fdt,
                                     device=self.device)
        flux = self._flux_struct(with_dtaus)
        st = self._stream()
        hist_ptr = None if T_hist is None else T_hist.data_ptr()
        prep_first = 1 if self._records_stale else 0
        if self.group is None and self.sweep_events is None:
            _cabi.check(self.lib.frei_b200_sweep_step(
                C.byref(self._tab), C.byref(self._spec), C.byref(self._atm), C.byref(flux),
                direction, float(alpha_override), C.byref(self._ws), hist_ptr, prep_first, 1, st))
            self.launches += 2 + prep_first
            self._records_stale = False
            return
        if prep_first:
            _cabi.check(self.lib.frei_b200_layer_prep(C.byref(self._tab), C.byref(self._atm),
                                                      C.byref(self._ws), None, None, None, None,
                                                      None, st))
        if self.sweep_events is not None:
            torch = _torch()
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _cabi.check(self.lib.frei_b200_sweep(C.byref(self._tab), C.byref(self._spec),
                                             C.byref(self._atm), C.byref(flux), direction,
                                             C.byref(self._ws), st))
        if self.sweep_events is not None:
            ev[1].record()
            self.sweep_events.append(ev)
        if self.group is None:
            _cabi.check(self.lib.frei_b200_post(C.byref(self._tab), C.byref(self._atm),
                                                C.byref(self._ws), self.n_lam, direction,
                                                float(alpha_override), hist_ptr, 1, st))
            self.launches += 2 + prep_first
        elif self._p2p is not None:
            p = self._p2p
            p['epoch'] += 1
            if p['epoch'] & 0xffffffff == 0:                   # the low word is the in-band flag: never 0
                p['epoch'] += 1
            arg = _cabi.frei_p2p(p['bufs'].data_ptr(), p['err'].data_ptr(), p['epoch'], p['rank'],
                                 p['world'])
            _cabi.check(self.lib.frei_b200_post_p2p(C.byref(self._tab), C.byref(self._atm),
                                                    C.byref(self._ws), self.n_lam, direction,
                                                    float(alpha_override), hist_ptr, 1,
                                                    C.byref(arg), st))
            self.launches += 2 + prep_first
        else:
            _cabi.check(self.lib.frei_b200_reduce(C.byref(self._atm), C.byref(self._ws),
                                                  self.n_lam, st))
            allreduce_sums(self.sums, self.group)
            _cabi.check(self.lib.frei_b200_update_T(C.byref(self._tab), C.byref(self._atm),
                                                    C.byref(self._ws), direction,
                                                    float(alpha_override), hist_ptr, st))
            self.launches += 3 + prep_first
        self._records_stale = False

    def emit(self, **kw):
        self.sweep(FREI_EMIT, **kw)

    def absorb(self, **kw):
        self.sweep(FREI_ABSORB, **kw)
