"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py:
per-kernel launches, mean duration and share of the time spent in this library's kernels.
usage: python scripts/launch_list_summary.py LAUNCHES.csv OUT.json"""
import csv
import json
import sys

OURS = ('sweep_kernel', 'sweep_f32_kernel', 'post_kernel', 'post_level_kernel', 'prep_kernel', 'spectral_kernel',
        'update_prep_kernel', 'kappa_kernel', 'propagate_kernel', 'diag_kernel', 'diag_finish_kernel',
        'bin_trapz', 'regrid_kernel')   # dfma_peak_kernel is the roofline's own measurement, not the path
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iname, ival = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = {}
for r in rows[1:]:
    k = r[iname]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ival].replace(',', ''))
ours = {k: v for k, v in agg.items() if any(o in k for o in OURS)}
tot = sum(v[1] for v in ours.values())
out = {
    'command': 'ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py '
               '--steps 2 --warmup 3 --no-extras --no-cpu-baseline',
    'note': 'cold-cache, serialised launches: compare shares, not absolutes; the list also holds the '
            'torch kernels of table generation and the e2e leg',
    'share_of_frei_b200_kernels': {
        k[:80]: {'launches': v[0], 'mean_ns': v[1] / v[0], 'share': v[1] / tot}
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1])},
    'kernels': [{'kernel': k[:120], 'launches': v[0], 'total_ns': v[1]}
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])][:25],
}
json.dump(out, open(sys.argv[2], 'w'), indent=1)
for k, v in out['share_of_frei_b200_kernels'].items():
    print(f"{v['share']:.3f} {v['launches']:3d} x {v['mean_ns'] / 1e3:8.1f} us  {k}")
