#!/bin/bash
# round 2, GPU call 22: packed (FFMA2) fp32 sweep — parity tests of the fp32 mode, timing
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "fp32 or f32" 2>&1 | tail -4
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --nlam 200000 200002 800000 2>&1 | grep -E "^L |rror"
python scripts/size_scan.py --flux-dtype 32 --table-dtype 64 --nlam 200000 2>&1 | grep -E "^L |rror"
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --L 100 --S 8 --nlam 125000 1000000 2>&1 | grep -E "^L |rror"
} > gpurun_out/r02_run22.log 2>&1
cat gpurun_out/r02_run22.log
