#!/bin/bash
# round 2, GPU call 44: per-level kernel with uniform load batches and no spills — suite + A/B against the previous commit
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for pass in 1 2; do
bash scripts/ab_bench.sh "--steps 20 --warmup 5 --no-extras" prev_post default
done
} > gpurun_out/r02_run44.log 2>&1
cat gpurun_out/r02_run44.log
