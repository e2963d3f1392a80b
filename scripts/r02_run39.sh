#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; echo "exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02g_bench_n2.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('value %.4e step %.4f' % (d['value'], d['ms_per_step']))
        c = d.get('c5'); print('c5', {k: c.get(k) for k in ('value', 'iterations', 'converged', 'ms_total', 'per_iteration_ms', 'collective', 'error')})
        print('c4', d['c4'].get('value'), 'c3', d['strong_c3'].get('value'))
PY
