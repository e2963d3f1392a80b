#!/bin/bash
# round 2, GPU call 21: fp32-arithmetic sweep — where does its time go (ncu full), V = 4 x 64 threads vs V = 2 x 128
mkdir -p gpurun_out
{
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --nlam 200000 200002 800000 800002 2>&1 | grep -E "^L |rror"
python scripts/prof_sweep.py --iters 3 --flux-dtype 32 --table-dtype 32 > gpurun_out/r02b_plain_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_f32 -s 4 -c 2 -f -o gpurun_out/r02b_c2_f32 \
    python scripts/prof_sweep.py --iters 3 --flux-dtype 32 --table-dtype 32 > gpurun_out/r02b_ncu_f32.log 2>&1
tail -2 gpurun_out/r02b_ncu_f32.log
} > gpurun_out/r02_run21.log 2>&1
cat gpurun_out/r02_run21.log
