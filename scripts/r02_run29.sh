#!/bin/bash
# round 2, GPU call 29: timing-only probe — two dependent layer responses per basic block (results wrong by construction)
mkdir -p gpurun_out
{
bash scripts/ab_libs.sh "--nlam 200000 800000" default norm3 dbl3 dbl4
} > gpurun_out/r02_run29.log 2>&1
cat gpurun_out/r02_run29.log
