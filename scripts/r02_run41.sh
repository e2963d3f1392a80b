#!/bin/bash
# round 2, GPU call 41 (2 GPUs): multi-GPU parity with the per-level kernel (p2p and nccl), bench lines at N = 2 and 1 with extras
mkdir -p gpurun_out
{
FREI_DIST_LOGDIR=gpurun_out/dist2d timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "n2 exit $?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "n1 exit $?"
python - <<'PY'
import json
for n in (1, 2):
    for ln in open('gpurun_out/r02h_bench_n%d.json' % n):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('N=%d value %.4e step %.4f ms kernel %.4f e2e %.3e' % (n, d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
            for k in ('strong_c3', 'fp32_c2', 'c4', 'c5'):
                if k in d: print('   ', k, d[k].get('value'), d[k].get('ms_per_step'), d[k].get('ms_total'), d[k].get('iterations'), d[k].get('error'))
PY
} > gpurun_out/r02_run41.log 2>&1
cat gpurun_out/r02_run41.log
