#!/bin/bash
# round 2, GPU calls 28 and 36 (8 GPUs): whole GPU suite incl. multi-GPU parity on 2 and 8 ranks, bench lines at N = 8, 4, 2, 1
mkdir -p gpurun_out
FREI_DIST_LOGDIR=gpurun_out/dist8c timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest8.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest8.log
tail -4 gpurun_out/r02e_pytest8.log
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02e_bench_n$n.json 2> gpurun_out/r02e_bench_n$n.err; echo "bench n$n exit $?"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench n1 exit $?"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    for ln in open('gpurun_out/r02e_bench_n%d.json' % n):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('N=%d value %.4e step %.4f ms kernel %.4f e2e %.3e' % (n, d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
            for k in ('strong_c3', 'fp32_c2', 'c4'):
                if k in d: print('   ', k, d[k].get('value'), d[k].get('ms_per_step'))
PY
