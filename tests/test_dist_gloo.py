"""
World-size-2 gloo test (CPU) of the wavelength-sharded mode: each rank sweeps its
wavelength slice (the per-slice arithmetic is played by the oracle here — the CUDA
kernels need a GPU), the [L][4] integrals are summed with the package's own
collective plumbing, and T, dT and the gathered spectrum must equal the
single-rank run.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import frei_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from frei_b200 import synthetic
    from frei_b200.sharding import shard_range, allreduce_sums, gather_lambda
    w = synthetic.make_workload(14, 301, 3)
    pl = w['planet']
    lo, hi = shard_range(w['n_lam'], rank, world)
    tabs = synthetic.host_tables(w, lam_index=np.arange(lo, hi))
    lam_cm = w['lam_um'] * 1e-4
    F = O.F_TOA(lam_cm, pl['T_star'], a_rstar=pl['a_rstar'])[lo:hi]
    wts = O.trapz_weights(lam_cm)[lo:hi]
    Fu, Fd = np.zeros((14, hi - lo)), np.zeros((14, hi - lo))
    T = w['T_init'].copy()
    for direction, fn in (('emit', O.emit), ('absorb', O.absorb), ('emit', O.emit)):
        out = fn(tabs, T, w['P_bar'], w['lam_um'][lo:hi], F, pl['g'], pl['m_bar'],
                 lambda a, b: w['mmr'][0], fluxes_up=Fu, fluxes_down=Fd, trapz_w=wts)
        sums = torch.from_numpy(out[6].copy())
        allreduce_sums(sums)
        dT = O.thermo_from_bol(sums.numpy(), T, w['P_bar'], pl['g'], pl['m_bar'], pl['alpha'],
                               direction)
        T = T - dT
    spec = gather_lambda(torch.from_numpy(Fu[-1].copy()), w['n_lam'])
    full = gather_lambda(torch.from_numpy(Fd.copy()), w['n_lam'])
    if rank == 0:
        q.put((T, dT, spec, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_lambda_sharded_iteration_equals_single_rank():
    from frei_b200 import synthetic
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    T2, dT2, spec2, Fd2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synthetic.make_workload(14, 301, 3)
    pl = w['planet']
    tabs = synthetic.host_tables(w)
    F = O.F_TOA(w['lam_um'] * 1e-4, pl['T_star'], a_rstar=pl['a_rstar'])
    Fu, Fd = np.zeros((14, 301)), np.zeros((14, 301))
    T = w['T_init'].copy()
    for fn in (O.emit, O.absorb, O.emit):
        out = fn(tabs, T, w['P_bar'], w['lam_um'], F, pl['g'], pl['m_bar'],
                 lambda a, b: w['mmr'][0], fluxes_up=Fu, fluxes_down=Fd)
        T, dT = out[2], out[5]
    np.testing.assert_allclose(T2, T, rtol=0, atol=1e-7)
    np.testing.assert_allclose(dT2, dT, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(spec2, Fu[-1], rtol=1e-9)
    np.testing.assert_allclose(Fd2, Fd, rtol=1e-9, atol=1e-280)
