#!/bin/bash
# round 2, GPU call 33: post kernel with its constant loads and pressure-only math before the grid dependency wait
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_stamps.so python scripts/stamp_probe.py C2 2>&1 | tail -20
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_stamps.so python scripts/stamp_probe.py C3 125000 2>&1 | tail -20
python scripts/size_scan.py --nlam 5000 200000 800000 2>&1 | grep -E "^L |rror"
python scripts/size_scan.py --L 100 --S 8 --nlam 125000 2>&1 | grep -E "^L |rror"
bash scripts/ab_bench.sh "--steps 20 --warmup 5 --no-extras" default default
} > gpurun_out/r02_run33.log 2>&1
cat gpurun_out/r02_run33.log
