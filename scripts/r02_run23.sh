#!/bin/bash
# round 2, GPU call 23: packed fp32 sweep with compile-time species count — parity, guard bands, timing, instruction count
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x -k "fp32 or f32 or guard" 2>&1 | tail -4
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --nlam 200000 200002 800000 2>&1 | grep -E "^L |rror"
python scripts/size_scan.py --flux-dtype 32 --table-dtype 32 --L 100 --S 8 --nlam 125000 1000000 2>&1 | grep -E "^L |rror"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.per_cycle_active --clock-control none -k regex:sweep_f32 -s 4 -c 2 python scripts/prof_sweep.py --iters 3 --flux-dtype 32 --table-dtype 32 2>&1 | grep -E "sweep_f32|inst_executed|duration|issue_active|dram__|warps_active"
} > gpurun_out/r02_run23.log 2>&1
cat gpurun_out/r02_run23.log
