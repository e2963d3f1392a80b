#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "exit $?"
tail -3 gpurun_out/r02g_bench_n1.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02g_bench_n1.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('value %.4e step %.4f' % (d['value'], d['ms_per_step']))
        print('c5', d.get('c5'))
PY
