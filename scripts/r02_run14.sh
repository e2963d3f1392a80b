#!/bin/bash
# round 2, GPU call 14: relay plan vs whole-chunk plans on the same box (A/B, two passes)
mkdir -p gpurun_out
{
for pass in 1 2; do
bash scripts/ab_libs.sh "--nlam 37888 100000 151552 200000 800000" default norelay
done
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000 1000000" default norelay
} > gpurun_out/r02_run14.log 2>&1
cat gpurun_out/r02_run14.log
