#!/bin/bash
# round 2, GPU call 31: timing-only removal probes of the single-layer kernel (results wrong by construction)
mkdir -p gpurun_out
{
bash scripts/ab_libs.sh "--nlam 37888 200000 800000" nopair x_nostore x_noload x_nostage x_nomem x_noreduce x_nogather x_nomath
} > gpurun_out/r02_run31.log 2>&1
cat gpurun_out/r02_run31.log
