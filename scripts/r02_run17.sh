#!/bin/bash
# round 2, GPU call 17: gather-before-reduction variant (A/B, two passes)
mkdir -p gpurun_out
{
for pass in 1 2; do
bash scripts/ab_libs.sh "--nlam 37888 151552 200000 800000" default gfirst
done
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000 1000000" default gfirst
} > gpurun_out/r02_run17.log 2>&1
cat gpurun_out/r02_run17.log
