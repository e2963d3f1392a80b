"""Summarise an `ncu --set full` report of the sweep kernel into profiles/ (run where ncu is installed).
usage: python scripts/ncu_summary.py REPORT.ncu-rep CAPTURE_TAG "what was captured" [--write [KEY [EVALS]]]
(EVALS = evaluations per captured launch, (L - 1) x n_lambda: lets bench.py scale the entry to its own launch size)
Prints one JSON object per captured launch; --write appends them to
profiles/r02_sweep_kernel_ncu_summary.json and refreshes the entry KEY (default C2_tab64_flux64_n1)
of profiles/traffic.json: DRAM bytes and executed fp64 thread operations per launch, which
bench.py turns into roofline.traffic and the fp64 flop count of its roofline."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
]
STALLS = ['dispatch_stall', 'long_scoreboard', 'math_pipe_throttle', 'not_selected', 'selected',
          'short_scoreboard', 'wait', 'barrier', 'lg_throttle', 'mio_throttle', 'no_instruction', 'branch_resolving']


def to_bytes(val, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(val.replace(',', '')) * scale.get(unit, 1)


def main():
    rep, tag, what = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(header)}
    res, traffic, ops = [], [], []
    for r in rows[2:]:
        d = {'capture': tag, 'what': what, 'Kernel Name': r[col['Kernel Name']]}
        for k in KEEP:
            if k in col:
                d[k] = f'{r[col[k]]} {units[col[k]]}'.strip()
        for s in STALLS:
            k = f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio'
            if k in col and r[col[k]]:
                d['stall_' + s] = round(float(r[col[k]].replace(',', '')), 2)
        rd, wr = 'dram__bytes_read.sum', 'dram__bytes_write.sum'
        if rd in col and wr in col:
            traffic.append(to_bytes(r[col[rd]], units[col[rd]]) + to_bytes(r[col[wr]], units[col[wr]]))
        # executed fp64 thread instructions of the launch: rate per elapsed cycle x elapsed cycles
        cyc = 'smsp__cycles_elapsed.avg'
        o = {}
        for op in ('dfma', 'dmul', 'dadd'):
            k = f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed'
            if k in col and cyc in col and r[col[k]]:
                o[op] = float(r[col[k]].replace(',', '')) * float(r[col[cyc]].replace(',', ''))
        if len(o) == 3:
            d['fp64_thread_ops'] = {k: round(v) for k, v in o.items()}
            ops.append(o)
        res.append(d)
    print(json.dumps(res, indent=1))
    if '--write' in sys.argv:
        p = os.path.join(ROOT, 'profiles', 'r02_sweep_kernel_ncu_summary.json')
        allc = (json.load(open(p)) if os.path.exists(p) else []) + res
        wi = sys.argv.index('--write')
        key = sys.argv[wi + 1] if len(sys.argv) > wi + 1 else 'C2_tab64_flux64_n1'
        evals = int(sys.argv[wi + 2]) if len(sys.argv) > wi + 2 else None
        json.dump(allc, open(p, 'w'), indent=1)
        if traffic:
            tp = os.path.join(ROOT, 'profiles', 'traffic.json')
            t = json.load(open(tp))
            t[key] = {
                'bytes_per_launch': sum(traffic) / len(traffic),
                'per_launch': traffic,
                'source': f'profiles/r02_sweep_kernel_ncu_summary.json capture {tag} '
                          f'(ncu --set full of scripts/prof_sweep.py)'}
            if evals:
                t[key]['evals_per_launch'] = evals
            if ops:      # thread-level fp64 instructions per launch (DFMA = 2 flops, DMUL / DADD = 1)
                t[key]['fp64_thread_ops_per_launch'] = {
                    k: sum(o[k] for o in ops) / len(ops) for k in ('dfma', 'dmul', 'dadd')}
            json.dump(t, open(tp, 'w'), indent=1)


if __name__ == '__main__':
    main()
