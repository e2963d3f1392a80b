#!/bin/bash
# round 2, GPU call 15: relay kernel as its own instantiation — suite, A/B against the whole-chunk build, bench
mkdir -p gpurun_out
{
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for pass in 1 2; do
bash scripts/ab_libs.sh "--nlam 37888 100000 151552 160000 200000 250000 303104 400000 800000" default norelay
done
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000 250000 1000000" default norelay
bash scripts/ab_libs.sh "--L 200 --S 3 --nlam 250000" default norelay
bash scripts/ab_bench.sh "--steps 20 --warmup 5 --no-extras" default norelay default norelay
} > gpurun_out/r02_run15.log 2>&1
cat gpurun_out/r02_run15.log
