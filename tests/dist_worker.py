"""
Worker of tests/test_gpu_dist.py (not collected by pytest):
    python -m torch.distributed.run --nproc-per-node N tests/dist_worker.py [--collective p2p|nccl|auto]
Wavelength-sharded Grid.emission_spectrum vs the single-GPU run of the same problem
(every rank computes both): spectrum, T history, dtaus must agree to rounding of the summation
order (<= 1e-9 relative on fluxes, 1e-5 K on temperatures after up to 40 iterations — the contract
is 0.1 K), same iteration count.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import frei_b200 as frei  # noqa: E402
from frei_b200 import synthetic  # noqa: E402
from frei_b200.opacity import OpacityTable  # noqa: E402

import argparse  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument('--collective', default='auto', choices=['auto', 'p2p', 'nccl'])
ap.add_argument('--nlam', type=int, default=30001)      # odd count: uneven shards, 32- and 64-wide chunks
ap.add_argument('--layers', type=int, default=40)
ap.add_argument('--species', type=int, default=3)
args = ap.parse_args()
rank = int(os.environ['RANK'])
local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))

w = synthetic.make_workload(args.layers, args.nlam, args.species)
tabs = synthetic.host_tables(w)
op = {k: OpacityTable(t['values'], t['P'], t['T'], w['lam_um']) for k, t in tabs.items()}
pl = w['planet']
planet = frei.Planet(a_rstar=pl['a_rstar'], m_bar=pl['m_bar'], g=pl['g'] / 100.0, T_star=pl['T_star'],
                     alpha=pl['alpha'])


def solve(group, gather='all'):
    grid = frei.Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
    grid.load_opacities(opacities=op)
    grid.collective = args.collective
    out = grid.emission_spectrum(n_timesteps=40, group=group, gather=gather)
    global used, last_grid
    used = 'p2p-fused' if grid.engine._p2p is not None else ('nccl' if group is not None else 'single')
    last_grid = grid
    return out, grid.n_iterations


(s1, T1, h1, d1), n1 = solve(None)
(s2, T2, h2, d2), n2 = solve(dist.group.WORLD)
rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-250)))
ok = (n1 == n2 and rel(s2.flux, s1.flux) < 1e-9 and np.abs(T2 - T1).max() < 1e-5
      and h1.shape == h2.shape and np.abs(h2 - h1).max() < 1e-5 and rel(d2, d1) < 1e-10)
print(f'rank {rank}: [{used}] iterations {n1}/{n2} spectrum rel {rel(s2.flux, s1.flux):.2e} '
      f'T {np.abs(T2 - T1).max():.2e} K dtaus {rel(d2, d1):.2e} -> {"OK" if ok else "MISMATCH"}', flush=True)
# gather='local': every rank returns its own wavelength slice; T_eff from the resident state
g1 = frei.Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
teff1 = float(frei.effective_temperature(g1, s1, d1, T1))
(s3, T3, h3, d3), n3 = solve(dist.group.WORLD, gather='local')
lo, hi = last_grid.lam_range
diag = last_grid.diagnostics(group=dist.group.WORLD)
ok3 = (n3 == n2 and np.array_equal(np.asarray(s3.flux), np.asarray(s2.flux)[lo:hi])
       and np.array_equal(d3, d2[:, lo:hi]) and np.array_equal(T3, T2)
       and np.asarray(s3.wavelength).shape[0] == hi - lo and abs(float(diag['T_eff']) - teff1) < 1e-6)
print(f'rank {rank}: gather=local slice [{lo}, {hi}) T_eff {float(diag["T_eff"]):.6f} vs {teff1:.6f} '
      f'-> {"OK" if ok3 else "MISMATCH"}', flush=True)
ok = ok and ok3
if args.collective != 'auto' and used != {'p2p': 'p2p-fused', 'nccl': 'nccl'}[args.collective]:
    print(f'rank {rank}: asked for {args.collective}, ran {used}', flush=True)
    ok = False
flag = torch.tensor([0 if ok else 1], device='cuda')
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item()))
