#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_sweep.py --iters 3 > gpurun_out/r02d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 4 -c 2 -f -o gpurun_out/r02d_post \
    python scripts/prof_sweep.py --iters 3 > gpurun_out/r02d_ncu_post.log 2>&1
tail -2 gpurun_out/r02d_ncu_post.log
