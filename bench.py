#!/usr/bin/env python
"""
bench.py — layer x wavelength-bin two-stream flux evaluations per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one radiative-equilibrium iteration (emit sweep + absorb sweep,
each with its (P,T) brackets, opacity gather, two-stream layer response,
wavelength integrals, cross-GPU sum and temperature update) of BASELINE.json
config C2: hot Jupiter, 50 layers x 200k wavelength bins, 3 opacity species,
synthetic seeded tables, fp64.  One iteration = 2 (L-1) n_lambda evaluations.
With N > 1 GPUs the wavelength axis is sharded, 200k bins per GPU (weak
scaling), with an NCCL all-reduce of the [L][4] integrals after every sweep.

--impl reference times the reference's CPU arithmetic (the numpy+scipy oracle:
the reference itself cannot be installed here or on the GPU box — astropy,
xarray, specutils, pyfastchem are absent and there is no network) on all host
cores, same workload, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'layer-lambda-bin two-stream flux evaluations/s'
UNIT = 'evals/s'
L_C2, NLAM_C2, S_C2, TREF_C2 = 50, 200_000, 3, 2400.0

# BASELINE.json configs: name -> (layers, wavelength bins, species, T_ref, scaling, text)
WORKLOADS = {
    'C2': (50, 200_000, 3, 2400.0, 'weak', 'C2: hot Jupiter, 50 layers x 200k lambda bins per GPU, 3 species '
           '(H2O+CO+CH4 synthetic tables), one RE iteration (emit+absorb) per step'),
    'C1': (50, 5_000, 3, 2400.0, 'weak', 'C1: hot Jupiter, 50 layers x 5k lambda bins, 3 species, one RE iteration per step'),
    'C3': (100, 1_000_000, 8, 3200.0, 'strong', 'C3: ultra-hot Jupiter, 100 layers x 1M lambda bins (global), '
           '8 species, lambda-sharded, one RE iteration per step'),
    'C5': (200, 2_000_000, 3, 2400.0, 'strong', 'C5: stress, 200 layers x 2M lambda bins (global), 3 species, '
           'fp64, lambda-sharded, one RE iteration per step'),
}


def read_traffic(workload, table_dtype, flux_dtype, world):
    """DRAM bytes per sweep launch from the committed ncu capture of this workload (or None)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            d = json.load(fh)
        e = d.get(f'{workload}_tab{table_dtype}_flux{flux_dtype}_n{world}')
        return (e['bytes_per_launch'], e['source']) if e else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


def read_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-i', str(self.index), '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                  'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(power) if power else None}


# ---------------------------------------------------------------------------
# CPU arms (the oracle is the checker/baseline, never the product)
# ---------------------------------------------------------------------------
def cpu_baseline_one_core(seconds_target=12.0):
    """Oracle on ONE core over a bounded sample of C2: 2 iterations on a 50k-bin slice."""
    from oracle.parallel import ParallelOracle
    n_sample = 50_000
    po = ParallelOracle((L_C2, NLAM_C2, S_C2, TREF_C2), n_workers=1, n_lam_sample=n_sample)
    po.iteration()                                   # warm-up (page in, caches)
    t0 = time.perf_counter()
    n_it = 0
    while n_it < 2 or (time.perf_counter() - t0 < seconds_target and n_it < 6):
        po.iteration()
        n_it += 1
    dt = time.perf_counter() - t0
    return {'value': po.evals_per_iteration * n_it / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': f'{n_it} RE iterations on the first {n_sample} of {NLAM_C2} bins of C2 '
                      f'(50 layers, 3 species), numpy+scipy oracle, 1 process'}


def run_reference(args):
    """--impl reference: oracle on all host cores, C2 grid, one iteration per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle.parallel import ParallelOracle, available_cores
    cores = available_cores()
    workers = max(1, min(cores, 64))
    n_lam = NLAM_C2 if workers >= 4 else 25_000 * workers
    po = ParallelOracle((L_C2, NLAM_C2, S_C2, TREF_C2), n_workers=workers, n_lam_sample=n_lam)
    for _ in range(args.warmup):
        po.iteration()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.iteration()
    dt = time.perf_counter() - t0
    po.close()
    value = po.evals_per_iteration * args.steps / dt
    sample = (f'one RE iteration per step on {n_lam} of {NLAM_C2} bins of C2, wavelength axis '
              f'split over {workers} processes (numpy+scipy oracle of the reference arithmetic)')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOADS['C2'][5], 'n_layers': L_C2,
                   'n_lambda_global': NLAM_C2 * max(1, args.gpus), 'n_species': S_C2,
                   'table_dtype': 'f64', 'flux_dtype': 'f64',
                   'parallelism': f'{workers} host processes, wavelength-sharded',
                   'bins_timed': n_lam},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': workers, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from frei_b200 import synthetic
    from frei_b200.core import Grid, Planet
    from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F32, FREI_F64, shard_range

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        group = dist.group.WORLD
    if args.gpus != world and rank == 0:
        print(f'warning: --gpus {args.gpus} but WORLD_SIZE={world}', file=sys.stderr)

    L, n_lam_w, S, T_ref, scaling, wl_text = WORKLOADS[args.workload]
    n_lam_global = n_lam_w * world if scaling == 'weak' else n_lam_w   # weak: bins per GPU fixed
    if args.flux_dtype == 32:
        args.table_dtype = 32
    tdtype = FREI_F32 if args.table_dtype == 32 else FREI_F64
    fdtype = FREI_F32 if args.flux_dtype == 32 else FREI_F64
    b_flux = 4 if fdtype == FREI_F32 else 8
    w = synthetic.make_workload(L, n_lam_global, S, T_ref, table_f32=(tdtype == FREI_F32))
    lo, hi = shard_range(n_lam_global, rank, world)
    table = synthetic.device_table(w, tdtype, lam_range=(lo, hi), device=dev)
    pl = w['planet']
    eng = Engine(table, w['lam_um'], w['P_bar'], w['T_init'], w['mmr'], g=pl['g'],
                 m_bar=pl['m_bar'], alpha=pl['alpha'], T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                 group=group, flux_dtype=fdtype, collective=args.collective)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.sweep(FREI_EMIT)
        eng.sweep(FREI_ABSORB)

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # The sweep kernel is timed live inside the timed region with CUDA events on its stream, on
    # every 4th step (events between two kernels cost ~3 us each and keep the next kernel from
    # queueing up behind the previous one, so bracketing every launch would slow the step it measures).
    events = []
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    ev0.record()
    for i in range(args.steps):
        eng.sweep_events = events if i % 4 == 0 else None
        step()
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    sweep_ms = [a.elapsed_time(b) for a, b in events]
    eng.sweep_events = None
    launches = eng.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    evals_per_step = 2 * (L - 1) * n_lam_global
    value = evals_per_step * args.steps / (ms * 1e-3)

    # roofline of the dominant kernel (the layer sweep), this rank's launches
    b_tab = 4 if tdtype == FREI_F32 else 8
    bytes_per_eval = 4 * S * b_tab + 3 * b_flux
    sweep_avg_ms = float(np.mean(sweep_ms))
    algo_bytes = (L - 1) * (hi - lo) * bytes_per_eval
    peak, peak_src = read_peaks()
    achieved = algo_bytes / (sweep_avg_ms * 1e-3) / 1e9
    traffic, traffic_src = read_traffic(args.workload, args.table_dtype, args.flux_dtype, world)
    # the pipe that actually limits the fp64 kernel: executed-path SASS counts of the
    # layer loop for S = 3, two wavelengths per thread, E = 1 form (scripts/sass_loop_mix.py):
    # 110 DFMA + 66 DMUL + 39 DADD per warp and layer = 64 evaluations
    fp64_info = None
    if S == 3 and fdtype == FREI_F64:
        sm_hz = 1.965e9
        evals_s = (L - 1) * (hi - lo) / (sweep_avg_ms * 1e-3)
        inst_per_eval, flops_per_eval = 215 / 64, (2 * 110 + 66 + 39) / 2
        # pipe cycles per warp instruction: 2, or 3 with three distinct register operands (60 of the
        # 110 DFMAs; measured, scripts/fp64_operands.cu) -> 490 cycles per warp and layer
        pipe_cycles_per_eval = (2 * 155 + 3 * 60) / 64
        fp64_info = {'warp_inst_per_eval': inst_per_eval, 'flops_per_eval': flops_per_eval,
                     'achieved_tflops': evals_s * flops_per_eval / 1e12,
                     'peak_tflops_nominal': 148 * 64 * 2 * sm_hz / 1e12,
                     'pipe_frac': evals_s * pipe_cycles_per_eval / (148 * 4 * sm_hz),
                     'note': 'fp64 pipe occupancy implied by the kernel time: pipe cycles of the '
                             'issued warp instructions / (148 SMs x 4 sub-partitions x 1.965 GHz); '
                             'ncu sm__pipe_fp64_cycles_active: 42-46 % (capture r1i)'}

    # e2e through the public API: Grid.emission_spectrum with host buffers
    e2e = None
    try:
        planet = Planet(a_rstar=pl['a_rstar'], m_bar=pl['m_bar'], g=pl['g'] / 100.0,
                        T_star=pl['T_star'], alpha=pl['alpha'])
        grid = Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
        grid.attach_device_table(table, species=w['species'])
        grid.flux_dtype = fdtype
        k_e2e = max(2, args.steps)
        # N > 1: every rank keeps its own wavelength slice of the results (gather='local'), so the
        # job's outputs cross N PCIe links in parallel; gather='all' would copy the N-times larger
        # weak-scaled result to the host of every rank
        grid.emission_spectrum(n_timesteps=2, n_zero_crossings=10 ** 9, convergence_dT=0, group=group,
                               gather='local')
        sync()
        t0 = time.perf_counter()
        grid.emission_spectrum(n_timesteps=k_e2e, n_zero_crossings=10 ** 9, convergence_dT=0,
                               group=group, gather='local')
        sync()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        n_loc = hi - lo
        h2d = (8 * n_lam_global + 8 * L * (2 + S) + 8 * 3) / k_e2e
        d2h = 2 * L * 8 + 0.25 + ((L + 1) * n_loc * 8 + L * 8) / k_e2e  # T history, flag polls, fp64 results
        e2e = {'value': (2 * k_e2e + 1) * (L - 1) * n_lam_global / dt, 'unit': UNIT,
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'call': f'Grid.emission_spectrum(n_timesteps={k_e2e}'
                       + (", group=WORLD, gather='local'" if world > 1 else '') +
                       ') incl. setup, device-side convergence rule polled every 4 iterations, '
                       'final emit, T history + spectrum + dtaus D2H into pinned host arrays'
                       + (' (each rank: its own wavelength slice)' if world > 1 else ''),
               'seconds': dt}
    except Exception as exc:                                   # pragma: no cover
        e2e = {'value': None, 'error': repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == 'C2':
        cpu = cpu_baseline_one_core()

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': scaling, 'vs_baseline': None, 'dtype': f'f{args.flux_dtype}', 'data': 'synthetic',
            'config': {
                'workload': wl_text,
                'n_layers': L, 'n_lambda_global': n_lam_global, 'n_species': S,
                'table_dtype': f'f{args.table_dtype}', 'flux_dtype': f'f{args.flux_dtype}',
                'parallelism': f'lambda-sharded x{world}' if world > 1 else 'single GPU',
                'collective': ('p2p-fused' if eng._p2p is not None else 'nccl') if world > 1 else None,
                'l2': 'inputs larger than L2: flux state '
                      f'{2 * L * (hi - lo) * b_flux / 1e6:.0f} MB + table '
                      f'{table.values.numel() * b_tab / 1e6:.0f} MB per GPU touched every step'
                      if 2 * L * (hi - lo) * b_flux + table.values.numel() * b_tab > 130e6 else
                      'working set fits L2 (flux state '
                      f'{2 * L * (hi - lo) * 8 / 1e6:.0f} MB): 256 MB scratch written between steps '
                      'outside the event pairs is NOT done; kernel times are warm-L2',
            },
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'traffic_source': traffic_src,
                         'dram_gbs': (traffic / (sweep_avg_ms * 1e-3) / 1e9) if traffic else None,
                         'note': 'achieved = algorithmic bytes (SURVEY 8d: 4 S table rows + 3 flux '
                                 'words per evaluation) / kernel time; table rows shared by the '
                                 'levels of one (P,T) cell are served on chip, so the DRAM traffic '
                                 '(traffic, dram_gbs) is lower and frac can exceed 1; the kernel '
                                 'is bound by fp64 issue + shared-memory bandwidth (DESIGN.md 3.1)',
                         'kernel': 'sweep_kernel', 'kernel_avg_ms': sweep_avg_ms,
                         'kernel_launches_timed': len(sweep_ms),
                         'bytes_per_eval': bytes_per_eval,
                         'kernel_share_of_step': 2 * sweep_avg_ms / (ms / args.steps),
                         'fp64': fp64_info},
            'cpu_baseline': cpu,
            'e2e': e2e,
            'gpu_launches': launches,
            'clocks': clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_batch(args):
    """
    C4: a grid of atmospheres (T_eq x log g x metallicity), 50 layers x 20k bins, 3 species, each
    iterated to convergence (Grid.emission_spectrum's rule, evaluated on the device) and finished
    with the final emit; atmospheres are sharded over the GPUs with no collective.
    Metric: converged T-P profiles per second.
    """
    import torch
    import torch.distributed as dist
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64, shard_range

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L, n_lam, S = 50, 20_000, 3
    n_side = max(1, round(args.batch ** (1 / 3)))
    B_total = n_side ** 3
    w = synthetic.make_workload(L, n_lam, S, 2400.0)
    pl = w['planet']
    # T_eq 1000..2500 K (sets the initial profile and the irradiation), log g 2.5..4.0 (cgs),
    # metallicity -1..+2 dex as a multiplier on the mixing ratios
    T_ref = np.linspace(1000, 2500, n_side)
    logg = np.linspace(2.5, 4.0, n_side)
    met = np.linspace(-1, 2, n_side)
    tt, gg, mm = [x.ravel() for x in np.meshgrid(T_ref, logg, met, indexing='ij')]
    # interleaved sharding: neighbouring grid points (similar iteration counts) go to different GPUs
    sel = slice(rank, B_total, world)
    Bl = len(range(rank, B_total, world))
    T0 = tt[sel, None] * (w['P_bar'][None, :] / 0.1) ** 0.1
    mmr = w['mmr'][None] * (10.0 ** mm[sel])[:, None, None]
    table = synthetic.device_table(w, FREI_F64, device=dev)
    eng = Engine(table, w['lam_um'], np.broadcast_to(w['P_bar'], (Bl, L)), T0, mmr, g=10.0 ** gg[sel],
                 m_bar=pl['m_bar'], alpha=1.0, T_star=pl['T_star'], a_rstar=pl['a_rstar'],
                 ftoa_scale=(tt[sel] / 2400.0) ** 4)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up on the real state, then reset
    for _ in range(max(3, args.warmup)):
        eng.iteration()
    eng.reset(T0, mmr)
    sampler = ClockSampler(local_rank)
    sync()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launches
    sync()
    ev0.record()
    iters, T = eng.solve_batch(args.max_iterations, check_every=8)
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    it_sum = torch.tensor([float(iters.sum()), float(iters.max()), float((iters >= args.max_iterations).sum())],
                          dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        agg = [torch.zeros_like(it_sum) for _ in range(world)]
        dist.all_gather(agg, it_sum)
        it_sum = torch.stack(agg)
        tot_it, max_it, capped = float(it_sum[:, 0].sum()), float(it_sum[:, 1].max()), float(it_sum[:, 2].sum())
    else:
        tot_it, max_it, capped = [float(x) for x in it_sum]
    ms = float(t.item())
    if rank == 0:
        evals = (2 * tot_it + B_total) * (L - 1) * n_lam          # sweeps actually executed
        print(json.dumps({
            'metric': 'converged T-P profiles/s', 'value': B_total / (ms * 1e-3), 'unit': 'profiles/s',
            'n_gpus': world, 'steps': 1, 'warmup': max(3, args.warmup), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': f'C4: batch of {B_total} atmospheres (T_eq x log g x metallicity), '
                                   '50 layers x 20k lambda bins, 3 species, full RE solve each, '
                                   'batch-sharded (no collective)',
                       'mean_iterations': tot_it / B_total, 'max_iterations': max_it,
                       'hit_iteration_cap': capped, 'iteration_cap': args.max_iterations,
                       'useful_evals_per_s': evals / (ms * 1e-3)},
            'gpu_launches': eng.launches - launches0, 'clocks': clocks}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--max-iterations', type=int, default=400)
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--table-dtype', type=int, default=64, choices=[32, 64])
    ap.add_argument('--flux-dtype', type=int, default=64, choices=[32, 64],
                    help='64: fp64 arithmetic (headline); 32: fp32 state and arithmetic, fp64 integrals')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--collective', default='auto', choices=['auto', 'p2p', 'nccl'],
                    help='N > 1: fused peer-memory all-reduce inside the post kernel (p2p) or NCCL')
    ap.add_argument('--workload', default='C2', choices=['C1', 'C2', 'C3', 'C4', 'C5'])
    ap.add_argument('--batch', type=int, default=4096, help='C4: atmospheres in total')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'C4':
        run_batch(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
