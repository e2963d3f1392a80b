#!/bin/bash
# round 2, GPU call 42 (4 GPUs): multi-GPU parity on 2 and 4 ranks and the 4-GPU bench line with the per-level kernel
mkdir -p gpurun_out
{
FREI_DIST_LOGDIR=gpurun_out/dist4d timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n4.json 2> gpurun_out/r02h_bench_n4.err; echo "n4 exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02h_bench_n4.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N=4 value %.4e step %.4f ms kernel %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
        for k in ('strong_c3', 'fp32_c2', 'c4', 'c5'):
            if k in d: print('   ', k, d[k].get('value'), d[k].get('ms_per_step'), d[k].get('ms_total'), d[k].get('iterations'), d[k].get('error'))
PY
} > gpurun_out/r02_run42.log 2>&1
cat gpurun_out/r02_run42.log
