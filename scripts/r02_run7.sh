#!/bin/bash
# round 2, GPU call 7 (8 GPUs): parity suite incl. multi-GPU tests on 2 and 8 ranks, bench at N = 8, 4 (the N = 1, 2 lines exist)
mkdir -p gpurun_out
FREI_DIST_LOGDIR=gpurun_out/dist8 timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest7.log
tail -5 gpurun_out/r02_pytest7.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "bench n$n exit $?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --collective nccl --no-extras > gpurun_out/r02_bench_n8_nccl.json 2> gpurun_out/r02_bench_n8_nccl.err; echo "bench n8 nccl exit $?"
