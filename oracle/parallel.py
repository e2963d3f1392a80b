"""
CPU ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY (see frei_oracle.py).

All-host-cores driver of the oracle for ``bench.py --impl reference`` and the
``cpu_baseline`` leg: the wavelength axis is split over worker processes, each
running the oracle's emit/absorb on its slice (numpy elementwise code is
single-threaded, exactly like the reference's); the parent sums the four
wavelength integrals per layer and applies the scalar temperature update.  It is
the same decomposition the GPU path uses across devices, so it is the most
favourable way to run the reference's arithmetic on every core.
"""
import multiprocessing as mp
import os

import numpy as np

from . import frei_oracle as O


def available_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:          # pragma: no cover
        return os.cpu_count() or 1


def _shard(n, r, world):
    base, rem = divmod(n, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


class Slice:
    """One wavelength slice of a synthetic workload, advanced sweep by sweep."""

    def __init__(self, wl_args, lo, hi, lam_stride=1):
        from frei_b200 import synthetic          # input generation only (no hot-path code)
        w = synthetic.make_workload(*wl_args)
        idx = np.arange(w['n_lam'])[::lam_stride][lo:hi]
        self.w = w
        self.tabs = synthetic.host_tables(w, lam_index=idx)
        self.lam_um = w['lam_um'][idx]
        lam_cm_all = w['lam_um'][::lam_stride] * 1e-4
        self.wts = O.trapz_weights(lam_cm_all)[lo:hi]
        pl = w['planet']
        self.F_toa = O.F_TOA(self.lam_um * 1e-4, pl['T_star'], a_rstar=pl['a_rstar'])
        self.Fu = np.zeros((w['L'], hi - lo))
        self.Fd = np.zeros((w['L'], hi - lo))
        self.mmr = w['mmr'][0]

    def sweep(self, direction, T):
        w, pl = self.w, self.w['planet']
        fn = O.emit if direction == 'emit' else O.absorb
        out = fn(self.tabs, T, w['P_bar'], self.lam_um, self.F_toa, pl['g'], pl['m_bar'],
                 lambda a, b: self.mmr, alpha=pl['alpha'], fluxes_up=self.Fu,
                 fluxes_down=self.Fd, trapz_w=self.wts)
        return out[6]


def _single_threaded_blas():
    """np.dot (the sharded trapezoid sums) would otherwise start one BLAS thread per core in
    every worker process; with N workers that oversubscribes the host N-fold (measured: 20x slower)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:                # pragma: no cover
        import contextlib
        return contextlib.nullcontext()


def _worker(conn, wl_args, lo, hi, lam_stride):
    with _single_threaded_blas():
        sl = Slice(wl_args, lo, hi, lam_stride)
        conn.send('ready')
        while True:
            msg = conn.recv()
            if msg is None:
                break
            direction, T = msg
            conn.send(sl.sweep(direction, T))


class ParallelOracle:
    """RE iterations of a synthetic workload on ``n_workers`` processes (1 = in-process)."""

    def __init__(self, wl_args, n_workers=1, lam_stride=1, n_lam_sample=None):
        from frei_b200 import synthetic
        self.w = synthetic.make_workload(*wl_args)
        n = len(self.w['lam_um'][::lam_stride])
        if n_lam_sample is not None:
            n = min(n, n_lam_sample)
        self.n_lam = n
        self.n_workers = n_workers
        self.T = self.w['T_init'].copy()
        self.procs, self.conns, self.local = [], [], None
        if n_workers == 1:
            self.local = Slice(wl_args, 0, n, lam_stride)
        else:
            ctx = mp.get_context('fork')
            for r in range(n_workers):
                lo, hi = _shard(n, r, n_workers)
                a, b = ctx.Pipe()
                p = ctx.Process(target=_worker, args=(b, wl_args, lo, hi, lam_stride), daemon=True)
                p.start()
                self.procs.append(p)
                self.conns.append(a)
            for c in self.conns:
                assert c.recv() == 'ready'

    @property
    def evals_per_iteration(self):
        return 2 * (self.w['L'] - 1) * self.n_lam

    def _sweep(self, direction):
        if self.local is not None:
            with _single_threaded_blas():        # "cores: 1" means one thread
                bol = self.local.sweep(direction, self.T)
        else:
            for c in self.conns:
                c.send((direction, self.T))
            bol = np.sum([c.recv() for c in self.conns], axis=0)
        pl = self.w['planet']
        dT = O.thermo_from_bol(bol, self.T, self.w['P_bar'], pl['g'], pl['m_bar'], pl['alpha'],
                               direction)
        self.T = self.T - dT
        return dT

    def iteration(self):
        self._sweep('emit')
        return self._sweep('absorb')

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
