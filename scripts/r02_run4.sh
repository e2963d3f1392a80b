#!/bin/bash
# round 2, GPU call 4 (2 GPUs): full parity suite incl. the multi-GPU tests, bench at N = 1 and N = 2
mkdir -p gpurun_out
FREI_DIST_LOGDIR=gpurun_out/dist timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest4.log
tail -6 gpurun_out/r02_pytest4.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench n1 exit $?"
tail -c 600 gpurun_out/r02_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 exit $?"
tail -c 600 gpurun_out/r02_bench_n2.err
