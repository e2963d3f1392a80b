#!/bin/bash
# round 2, GPU call 30: two layer-steps per loop iteration — parity / bit-identity / guard bands, then A/B
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
bash scripts/ab_libs.sh "--nlam 37888 100000 151552 200000 800000" nopair default pair3
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000 1000000" nopair default
} > gpurun_out/r02_run30.log 2>&1
cat gpurun_out/r02_run30.log
