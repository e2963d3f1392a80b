#!/bin/bash
# round 2, GPU call 13: relay plan — parity/memory-safety suite, size scan, bench
mkdir -p gpurun_out
{
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== size_scan relay"
timeout 300 python scripts/size_scan.py --nlam 37888 100000 151552 160000 200000 250000 303104 400000 800000 2>&1 | grep -E "^L |rror"
timeout 300 python scripts/size_scan.py --L 100 --S 8 --nlam 125000 250000 1000000 2>&1 | grep -E "^L |rror"
timeout 300 python scripts/size_scan.py --L 200 --S 3 --nlam 250000 2>&1 | grep -E "^L |rror"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('value %.4e  step %.4f ms  sweep %.4f ms  e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
"
} > gpurun_out/r02_run13.log 2>&1
cat gpurun_out/r02_run13.log
