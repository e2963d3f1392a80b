#!/bin/bash
# round 2, GPU call 37: batch mode keeps the pressure-only terms of K4 in the finishing CTA — suite + C4 record
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02f_bench_n1.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('value %.4e step %.4f kernel %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
        for k in ('strong_c3', 'fp32_c2', 'c4'):
            print(k, d[k].get('value'), d[k].get('ms_per_step'), d[k].get('ms_total'))
PY
} > gpurun_out/r02_run37.log 2>&1
cat gpurun_out/r02_run37.log
