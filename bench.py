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


def read_traffic(workload, table_dtype, flux_dtype, world, evals=None):
    """
    DRAM bytes per sweep launch from the committed ncu capture of this workload (or None).  With
    `evals` (evaluations of the launch in question) the captured bytes are scaled per evaluation,
    so the C3 entry, captured on one of eight shards, also serves the other shard sizes.
    """
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            d = json.load(fh)
        e = d.get(f'{workload}_tab{table_dtype}_flux{flux_dtype}_n{world}')
        if not e:
            return None, None
        b = e['bytes_per_launch']
        if evals and e.get('evals_per_launch'):
            b = b * evals / e['evals_per_launch']
        return b, e['source']
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
# fp64 thread-level instructions per evaluation of the sweep kernel, used when profiles/traffic.json
# has no ncu entry for the workload (values of the committed C2 capture, S = 3)
FALLBACK_OPS_PER_EVAL = {'dfma': 55.0, 'dmul': 33.1, 'dadd': 20.4}


def read_fp64_ops(workload, table_dtype, flux_dtype):
    """fp64 thread instructions per evaluation from the committed ncu capture of this workload."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            e = json.load(fh)[f'{workload}_tab{table_dtype}_flux{flux_dtype}_n1']
        ops, n = e['fp64_thread_ops_per_launch'], e['evals_per_launch']
        return {k: ops[k] / n for k in ('dfma', 'dmul', 'dadd')}, e['source']
    except (OSError, ValueError, KeyError):
        return None, None


def distinct_table_rows(w, S):
    """
    Number of distinct (species, P node, T node) table rows the levels of the initial profile
    touch: the compulsory table traffic of one sweep is that many rows of n_lambda elements
    (a row shared by neighbouring levels, or by neighbouring (P,T) cells, is read from HBM once
    if it stays on chip).  Same bracket rule as K0 (scipy find_indices).
    """
    def idx(x, v):
        i = np.searchsorted(x, v, side='right') - 1
        return np.clip(i, 0, len(x) - 2)
    ip, it = idx(w['axis_P'], w['P_bar']), idx(w['axis_T'], w['T_init'])
    rows = set()
    for a, b in zip(ip.tolist(), it.tolist()):
        rows.update([(a, b), (a, b + 1), (a + 1, b), (a + 1, b + 1)])
    return S * len(rows)


def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    group = None
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    if world > 1:
        group = dist.group.WORLD
    return world, rank, local_rank, dev, group


def time_sweeps(args, workload, world, rank, dev, group, tdtype, fdtype, steps, warmup, sample_clocks=False,
                local_rank=0):
    """
    One workload on this job's GPUs: W untimed + K timed RE iterations (emit + absorb), CUDA events,
    max over ranks.  Returns a dict with everything the JSON line reports about it.
    """
    import torch
    import torch.distributed as dist
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_EMIT, FREI_ABSORB, FREI_F32, shard_range
    L, n_lam_w, S, T_ref, scaling, wl_text = WORKLOADS[workload]
    n_lam_global = n_lam_w * world if scaling == 'weak' else n_lam_w   # weak: bins per GPU fixed
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

    for _ in range(warmup):
        step()
    sync()
    sampler = ClockSampler(local_rank)
    if sample_clocks and rank == 0:
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
    for i in range(steps):
        eng.sweep_events = events if i % 4 == 0 else None
        step()
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    sweep_ms = [a.elapsed_time(b) for a, b in events]
    eng.sweep_events = None
    eng.check_errors()
    clocks = sampler.stop() if (sample_clocks and rank == 0) else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    evals_per_step = 2 * (L - 1) * n_lam_global
    return dict(w=w, table=table, eng=eng, L=L, S=S, n_lam_global=n_lam_global, lo=lo, hi=hi,
                scaling=scaling, text=wl_text, ms=ms, steps=steps, warmup=warmup,
                value=evals_per_step * steps / (ms * 1e-3), sweep_avg_ms=float(np.mean(sweep_ms)),
                sweeps_timed=len(sweep_ms), launches=eng.launches - launches0, clocks=clocks,
                collective=('p2p-fused' if eng._p2p is not None else 'nccl') if world > 1 else None)


def fp64_peak(dev):
    """Thread-level DFMA/s measured now on this GPU (frei_b200_fp64_peak)."""
    import ctypes as C
    import torch
    from frei_b200 import _cabi
    lib = _cabi.load()
    n = 2048 * torch.cuda.get_device_properties(dev).multi_processor_count
    scratch = torch.empty(n, dtype=torch.float64, device=dev)
    out = C.c_double()
    _cabi.check(lib.frei_b200_fp64_peak(scratch.data_ptr(), n, C.byref(out),
                                        torch.cuda.current_stream(dev).cuda_stream))
    return out.value


def roofline_of(r, workload, tdtype, fdtype, dev):
    """Roofline of the sweep kernel of run `r` (workload name `workload`) against the resource that binds it."""
    from frei_b200.engine import FREI_F32
    tbits, fbits = (32 if tdtype == FREI_F32 else 64), (32 if fdtype == FREI_F32 else 64)
    L, S, n_loc = r['L'], r['S'], r['hi'] - r['lo']
    b_tab = 4 if tdtype == FREI_F32 else 8
    b_flux = 4 if fdtype == FREI_F32 else 8
    t = r['sweep_avg_ms'] * 1e-3
    evals = (L - 1) * n_loc
    peak_hbm, peak_src = read_peaks()
    algo_bytes = evals * (4 * S * b_tab + 3 * b_flux)
    reuse_bytes = evals * 3 * b_flux + distinct_table_rows(r['w'], S) * n_loc * b_tab
    traffic, traffic_src = read_traffic(workload, tbits, fbits, 1, evals)
    hbm = {'peak': peak_hbm, 'peak_source': peak_src, 'unit': 'GB/s',
           'algorithmic_bytes': algo_bytes, 'algorithmic_gbs': algo_bytes / t / 1e9,
           'reuse_aware_bytes': reuse_bytes, 'reuse_aware_gbs': reuse_bytes / t / 1e9,
           'reuse_aware_frac': reuse_bytes / t / 1e9 / peak_hbm,
           'traffic': traffic, 'traffic_source': traffic_src,
           'traffic_gbs': (traffic / t / 1e9) if traffic else None,
           'note': 'algorithmic = SURVEY 8d (4 S table rows + 3 flux words per evaluation, every level '
                   're-reads its rows); reuse_aware = 3 flux words per evaluation + every DISTINCT table '
                   'row the levels touch once (rows shared by the levels of a (P,T) cell stay on chip); '
                   'traffic = dram__bytes_read + write of one launch (ncu --set full capture)'}
    common = {'kernel': 'sweep_kernel', 'kernel_avg_ms': r['sweep_avg_ms'],
              'kernel_launches_timed': r['sweeps_timed'], 'evals_per_launch': evals,
              'kernel_share_of_step': 2 * r['sweep_avg_ms'] / (r['ms'] / r['steps']), 'hbm': hbm}
    if fdtype == FREI_F32:
        # fp32 arithmetic: the kernel streams; compulsory (reuse-aware) bytes against the copy bandwidth
        return dict(common, bound='hbm', achieved=hbm['reuse_aware_gbs'], peak=peak_hbm, unit='GB/s',
                    frac=hbm['reuse_aware_frac'], traffic=traffic,
                    note='fp32 arithmetic: bound by HBM; achieved = reuse-aware compulsory bytes / kernel '
                         'time (the SURVEY 8d algorithmic bytes would count table rows that never leave '
                         'the chip: see hbm.algorithmic_gbs)')
    # fp64 arithmetic: the fp64 pipe binds (ncu: 45 % pipe-active vs 42 % of HBM on real traffic)
    ops, ops_src = read_fp64_ops(workload, tbits, fbits)
    if ops is None:
        ops, ops_src = FALLBACK_OPS_PER_EVAL, 'fallback: counts of the committed C2 capture (S = 3)'
    dfma_peak = fp64_peak(dev)                                             # thread DFMA/s, measured now
    flops_per_eval = 2 * ops['dfma'] + ops['dmul'] + ops['dadd']
    slots_per_eval = ops['dfma'] + ops['dmul'] + ops['dadd']
    achieved = flops_per_eval * evals / t / 1e12
    peak = 2 * dfma_peak / 1e12
    return dict(common, bound='fp64', achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
                traffic=traffic,
                pipe_frac=slots_per_eval * evals / t / dfma_peak,
                fp64={'thread_ops_per_eval': ops, 'ops_source': ops_src, 'flops_per_eval': flops_per_eval,
                      'peak_dfma_per_s': dfma_peak,
                      'peak_source': 'measured in this run: frei_b200_fp64_peak (register-operand DFMA '
                                     'chains, 16 warps per scheduler, best of 3, CUDA events)'},
                note='fp64 arithmetic: bound by the fp64 pipe.  achieved = executed fp64 flops per launch '
                     '(ncu thread-level DFMA x 2 + DMUL + DADD of the committed capture, per evaluation) '
                     '/ kernel time; peak = DFMA rate measured in this run x 2.  pipe_frac counts issue '
                     'slots (a DMUL or DADD occupies the pipe as long as a DFMA).  The HBM view is under '
                     '"hbm": real DRAM traffic is ~1/3 of the SURVEY 8d algorithmic bytes.')


def run_e2e(args, r, world, rank, dev, group, fdtype):
    """The same metric through the public API with host buffers: Grid.emission_spectrum."""
    import torch
    import torch.distributed as dist
    from frei_b200.core import Grid, Planet
    w, L, S = r['w'], r['L'], r['S']
    pl = w['planet']

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    try:
        planet = Planet(a_rstar=pl['a_rstar'], m_bar=pl['m_bar'], g=pl['g'] / 100.0,
                        T_star=pl['T_star'], alpha=pl['alpha'])
        grid = Grid(planet, lam=w['lam_um'], pressures=w['P_bar'], init_temperatures=w['T_init'])
        grid.attach_device_table(r['table'], species=w['species'])
        grid.flux_dtype = fdtype
        k_e2e = max(2, args.steps)
        # N > 1: every rank keeps its own wavelength slice of the results (gather='local'), so the
        # job's outputs cross N PCIe links in parallel; gather='all' would copy the N-times larger
        # weak-scaled result to the host of every rank
        grid.emission_spectrum(n_timesteps=2, n_zero_crossings=10 ** 9, convergence_dT=0, group=group,
                               gather='local')
        sync()
        dts = []
        for _ in range(3):                                     # three whole solves, the median counts
            t0 = time.perf_counter()
            grid.emission_spectrum(n_timesteps=k_e2e, n_zero_crossings=10 ** 9, convergence_dT=0,
                                   group=group, gather='local')
            sync()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dts.append(float(tt.item()))
        dt = sorted(dts)[1]
        n_loc = r['hi'] - r['lo']
        # per solve: T, P, mmr, g, m_bar, alpha up (the wavelength grid and the opacity table are
        # resident, as in the reference where they are loaded once per Grid); down: T history
        # (2 L doubles per iteration), convergence flags, and once spectrum + dtaus + final T
        h2d = (8 * L * (2 + S) + 8 * 3) / k_e2e
        d2h = 2 * L * 8 + 0.25 + ((L + 1) * n_loc * 8 + L * 8) / k_e2e
        return {'value': (2 * k_e2e + 1) * (L - 1) * r['n_lam_global'] / dt, 'unit': UNIT,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'call': f'Grid.emission_spectrum(n_timesteps={k_e2e}'
                        + (", group=WORLD, gather='local'" if world > 1 else '') +
                        ') incl. per-solve setup, device-side convergence rule, final emit, '
                        'T history + spectrum + dtaus D2H into pinned host arrays'
                        + (' (each rank: its own wavelength slice)' if world > 1 else '')
                        + '; the opacity table upload is load-time (Grid.load_opacities) and outside',
                'seconds': dt, 'seconds_all': dts, 'frac_of_device_rate': None}
    except Exception as exc:                                   # pragma: no cover
        return {'value': None, 'error': repr(exc)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from frei_b200.engine import FREI_F32, FREI_F64

    world, rank, local_rank, dev, group = setup_dist()
    if args.gpus != world and rank == 0:
        print(f'warning: --gpus {args.gpus} but WORLD_SIZE={world}', file=sys.stderr)
    if args.flux_dtype == 32:
        args.table_dtype = 32
    tdtype = FREI_F32 if args.table_dtype == 32 else FREI_F64
    fdtype = FREI_F32 if args.flux_dtype == 32 else FREI_F64
    b_flux = 4 if fdtype == FREI_F32 else 8
    b_tab = 4 if tdtype == FREI_F32 else 8

    r = time_sweeps(args, args.workload, world, rank, dev, group, tdtype, fdtype, args.steps, args.warmup,
                    sample_clocks=True, local_rank=local_rank)
    L, S, lo, hi = r['L'], r['S'], r['lo'], r['hi']
    roof = roofline_of(r, args.workload, tdtype, fdtype, dev)
    e2e = run_e2e(args, r, world, rank, dev, group, fdtype)
    if e2e.get('value'):
        e2e['frac_of_device_rate'] = e2e['value'] / r['value']
    table_numel = r['table'].values.numel()
    for k in ('eng', 'table', 'w'):
        r[k] = None
    torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == 'C2':
        cpu = cpu_baseline_one_core()

    # secondary records: the other BASELINE.json configurations this job size is quoted on
    extras = {}
    if not args.no_extras and args.workload == 'C2' and args.flux_dtype == 64 and args.table_dtype == 64:
        try:
            c3 = time_sweeps(args, 'C3', world, rank, dev, group, FREI_F64, FREI_F64, steps=6, warmup=3)
            extras['strong_c3'] = {
                'metric': METRIC, 'value': c3['value'], 'unit': UNIT, 'scaling': 'strong',
                'ms_per_step': c3['ms'] / c3['steps'], 'steps': c3['steps'], 'warmup': c3['warmup'],
                'kernel_avg_ms': c3['sweep_avg_ms'], 'n_gpus': world, 'collective': c3['collective'],
                'workload': WORKLOADS['C3'][5], 'n_layers': c3['L'], 'n_lambda_global': c3['n_lam_global'],
                'n_species': c3['S'], 'dtype': 'f64', 'roofline': roofline_of(c3, 'C3', FREI_F64, FREI_F64, dev)}
            c3 = None
            torch.cuda.empty_cache()
        except Exception as exc:                               # pragma: no cover
            extras['strong_c3'] = {'error': repr(exc)}
        try:
            f32 = time_sweeps(args, 'C2', world, rank, dev, group, FREI_F32, FREI_F32, steps=args.steps,
                              warmup=3)
            extras['fp32_c2'] = {
                'metric': METRIC, 'value': f32['value'], 'unit': UNIT, 'scaling': f32['scaling'],
                'ms_per_step': f32['ms'] / f32['steps'], 'dtype': 'f32', 'n_gpus': world,
                'note': 'fp32 table, flux state and arithmetic, fp64 wavelength integrals (contract 1e-4)',
                'roofline': roofline_of(f32, 'C2', FREI_F32, FREI_F32, dev)}
            f32 = None
            torch.cuda.empty_cache()
        except Exception as exc:                               # pragma: no cover
            extras['fp32_c2'] = {'error': repr(exc)}
        try:
            extras['c4'] = batch_solve(args, world, rank, dev)
        except Exception as exc:                               # pragma: no cover
            extras['c4'] = {'error': repr(exc)}
        try:
            extras['c5'] = stress_solve(args, world, rank, dev, group)
        except Exception as exc:                               # pragma: no cover
            extras['c5'] = {'error': repr(exc)}

    if rank == 0:
        out = {
            'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': r['ms'] / args.steps, 'higher_is_better': True,
            'scaling': r['scaling'], 'vs_baseline': None, 'dtype': f'f{args.flux_dtype}', 'data': 'synthetic',
            'config': {
                'workload': r['text'],
                'n_layers': L, 'n_lambda_global': r['n_lam_global'], 'n_species': S,
                'table_dtype': f'f{args.table_dtype}', 'flux_dtype': f'f{args.flux_dtype}',
                'parallelism': f'lambda-sharded x{world}' if world > 1 else 'single GPU',
                'collective': r['collective'],
                'l2': 'inputs larger than L2: flux state '
                      f'{2 * L * (hi - lo) * b_flux / 1e6:.0f} MB + table '
                      f'{table_numel * b_tab / 1e6:.0f} MB per GPU touched every step'
                      if 2 * L * (hi - lo) * b_flux + table_numel * b_tab > 130e6 else
                      'working set fits L2 (flux state '
                      f'{2 * L * (hi - lo) * 8 / 1e6:.0f} MB): kernel times are warm-L2',
            },
            'roofline': roof,
            'cpu_baseline': cpu,
            'e2e': e2e,
            'gpu_launches': r['launches'],
            'clocks': r['clocks'],
        }
        out.update(extras)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def stress_solve(args, world, rank, dev, group, cap=400):
    """BASELINE config C5: 200 layers x 2M wavelength bins in fp64, wavelength-sharded over this job's
    GPUs, iterated to the reference's convergence rule (frei/core.py:301-318, on the device)."""
    import torch
    import torch.distributed as dist
    from frei_b200.engine import FREI_F64
    r = time_sweeps(args, 'C5', world, rank, dev, group, FREI_F64, FREI_F64, steps=2, warmup=2)
    eng, w, L = r['eng'], r['w'], r['L']
    eng.reset(w['T_init'])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launches
    ev0.record()
    iters, T = eng.solve_batch(cap, n_zero_crossings=2, convergence_dT=3.0, check_every=4)
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_it = int(iters[0])
    evals = (2 * n_it + 1) * (L - 1) * r['n_lam_global']       # the final emit included
    return {'metric': METRIC, 'value': evals / (ms * 1e-3), 'unit': UNIT, 'scaling': 'strong', 'n_gpus': world,
            'dtype': 'f64', 'workload': 'C5: stress, 200 layers x 2M lambda bins (global), 3 species, fp64, lambda-sharded, '
                                        'RE iteration to the convergence rule of frei/core.py:301-318 + final emit',
            'n_layers': L, 'n_lambda_global': r['n_lam_global'],
            'iterations': n_it, 'iteration_cap': cap, 'converged': bool(n_it < cap), 'ms_total': ms,
            'T_finite': bool(np.isfinite(T).all()), 'T_bottom_top_K': [float(T[0, 0]), float(T[0, -1])],
            'per_iteration_ms': r['ms'] / r['steps'], 'gpu_launches': eng.launches - l0,
            'collective': r['collective']}


def batch_solve(args, world, rank, dev):
    """
    C4: a grid of atmospheres (T_eq x log g x metallicity), 50 layers x 20k bins, 3 species, each
    iterated until Grid.emission_spectrum's rule (evaluated on the device) stops it and finished
    with the final emit; atmospheres are sharded over the GPUs with no collective.
    Returns the record of the metric "converged T-P profiles per second".
    """
    import torch
    import torch.distributed as dist
    from frei_b200 import synthetic
    from frei_b200.engine import Engine, FREI_F64

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
    for _ in range(3):
        eng.iteration()
    eng.reset(T0, mmr)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launches
    sync()
    ev0.record()
    iters, T = eng.solve_batch(args.max_iterations, check_every=8)
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    finite = np.isfinite(T).all(axis=1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(iters.sum()), float((iters >= args.max_iterations).sum()), float(finite.sum()),
                        float(eng.launches - launches0)], dtype=torch.float64, device=dev)
    mx = torch.tensor([float(iters.max())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    tot_it, capped, n_finite, launches = [float(x) for x in agg.tolist()]
    evals = (2 * tot_it + B_total) * (L - 1) * n_lam          # sweeps actually executed
    del eng, table
    torch.cuda.empty_cache()
    return {
        'metric': 'converged T-P profiles/s', 'value': n_finite / (ms * 1e-3), 'unit': 'profiles/s',
        'n_gpus': world, 'ms_total': ms, 'scaling': 'strong', 'dtype': 'f64',
        'workload': f'C4: batch of {B_total} atmospheres (T_eq x log g x metallicity), '
                    '50 layers x 20k lambda bins, 3 species, full RE solve each (until the convergence '
                    'rule of frei/core.py:306-318 stops it, then the final emit), batch-sharded, no collective',
        'atmospheres': B_total, 'converged_finite': int(n_finite),
        'diverged_nan': int(B_total - n_finite),
        'hit_iteration_cap': int(capped), 'iteration_cap': args.max_iterations,
        'mean_iterations': tot_it / B_total, 'max_iterations': float(mx.item()),
        'all_profiles_per_s': B_total / (ms * 1e-3), 'useful_evals_per_s': evals / (ms * 1e-3),
        'gpu_launches': int(launches),
        'note': 'value counts only atmospheres that end with a finite T-P profile; diverged_nan are cold, '
                'metal-rich corners of the grid where the reference\'s explicit temperature update '
                'overshoots to T < 0 and then NaN (the CPU oracle does the same); the reference\'s rule '
                'stops them too, because np.sign(nan) != np.sign(nan) counts as a zero crossing'}


def run_batch(args):
    """--workload C4: the batch record as the JSON line."""
    import torch.distributed as dist
    world, rank, local_rank, dev, group = setup_dist()
    rec = batch_solve(args, world, rank, dev)
    if rank == 0:
        out = {'metric': rec['metric'], 'value': rec['value'], 'unit': rec['unit'], 'n_gpus': world,
               'steps': 1, 'warmup': 3, 'ms_per_step': rec['ms_total'], 'higher_is_better': True,
               'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
               'config': {k: rec[k] for k in ('workload', 'atmospheres', 'converged_finite', 'diverged_nan',
                                              'hit_iteration_cap', 'iteration_cap', 'mean_iterations',
                                              'max_iterations', 'all_profiles_per_s', 'useful_evals_per_s',
                                              'note')},
               'gpu_launches': rec['gpu_launches']}
        print(json.dumps(out))
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
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the secondary records (C3 strong scaling, C2 in fp32, C4 batch solve)')
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
