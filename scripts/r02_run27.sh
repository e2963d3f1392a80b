#!/bin/bash
mkdir -p gpurun_out
{
for pass in 1 2; do
bash scripts/ab_bench.sh "--steps 20 --warmup 5 --flux-dtype 32 --table-dtype 32 --no-extras" default f32nopdl f32notrig
done
} > gpurun_out/r02_run27.log 2>&1
cat gpurun_out/r02_run27.log
