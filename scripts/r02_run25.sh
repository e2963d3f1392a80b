#!/bin/bash
# round 2, GPU call 25: packed fp32 sweep (final form) — full GPU suite, fp32 timing, bench secondary records
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --nlam 5000 200000 200002 800000 2>&1 | grep -E "^L |rror"
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --L 100 --S 8 --nlam 125000 1000000 2>&1 | grep -E "^L |rror"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02c_bench_n1.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('value %.4e step %.4f kernel %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
        for k in ('strong_c3', 'fp32_c2', 'c4'):
            print(k, d[k].get('value'), d[k].get('ms_per_step'), (d[k].get('roofline') or {}).get('kernel_avg_ms'), (d[k].get('roofline') or {}).get('frac'))
PY
} > gpurun_out/r02_run25.log 2>&1
cat gpurun_out/r02_run25.log
