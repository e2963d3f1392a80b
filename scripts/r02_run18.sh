#!/bin/bash
# round 2, GPU call 18 (2 GPUs): multi-GPU parity tests and the 2-GPU bench line with the relay kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02b_bench_n2.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N=2 value %.4e step %.4f ms kernel %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
        for k in ('strong_c3', 'fp32_c2', 'c4'):
            if k in d: print(k, d[k].get('value'), d[k].get('ms_per_step'))
PY
