#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -q -x -k "fp32 or f32 or guard" 2>&1 | tail -2
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --nlam 200000 800000 2>&1 | grep -E "^L |rror"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --flux-dtype 32 --table-dtype 32 --no-extras | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('fp32 value %.4e  step %.4f ms  sweep %.4f ms  e2e %.3e frac %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value'], d['roofline']['frac']))
"
} > gpurun_out/r02_run26.log 2>&1
cat gpurun_out/r02_run26.log
