#!/bin/bash
# round 2, GPU call 2: full parity suite with the unified sweep kernel, then ncu (full set, source
# counters) of the sweep with one warp per scheduler (37888 bins, 64-wide chunks) and of C2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest2.log
tail -5 gpurun_out/r02_pytest2.log
python scripts/prof_sweep.py --nlam 37888 --plan 2 --iters 3 > gpurun_out/r02_plain_lone.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/r02_lone \
    python scripts/prof_sweep.py --nlam 37888 --plan 2 --iters 3 > gpurun_out/r02_ncu_lone.log 2>&1
tail -2 gpurun_out/r02_ncu_lone.log
python scripts/prof_sweep.py --iters 3 > gpurun_out/r02_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 4 -c 2 -f -o gpurun_out/r02_c2 \
    python scripts/prof_sweep.py --iters 3 > gpurun_out/r02_ncu_c2.log 2>&1
tail -2 gpurun_out/r02_ncu_c2.log
ls -la gpurun_out/*.ncu-rep
