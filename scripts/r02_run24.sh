#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_sweep.py --iters 3 --flux-dtype 32 --table-dtype 32 > gpurun_out/r02b_plain_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_f32 -s 4 -c 2 -f -o gpurun_out/r02b_c2_f32_packed \
    python scripts/prof_sweep.py --iters 3 --flux-dtype 32 --table-dtype 32 > gpurun_out/r02b_ncu_f32.log 2>&1
tail -2 gpurun_out/r02b_ncu_f32.log
