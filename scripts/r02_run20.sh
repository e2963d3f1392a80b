#!/bin/bash
# round 2, GPU call 20: single-sync solve + cached optional imports + overlapped flag loads: suite, e2e phases, bench
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/e2e_phases.py 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('value %.4e  step %.4f ms  sweep %.4f ms  e2e %.3e  %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value'], d['e2e']['seconds_all']))
"
} > gpurun_out/r02_run20.log 2>&1
cat gpurun_out/r02_run20.log
