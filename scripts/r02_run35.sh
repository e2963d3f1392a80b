#!/bin/bash
# round 2, GPU call 35: finisher CTA + rehearsal of the serial tail in post_kernel
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_stamps.so python scripts/stamp_probe.py C2 2>&1 | tail -24
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_stamps.so python scripts/stamp_probe.py C3 125000 2>&1 | tail -24
for pass in 1 2; do
bash scripts/ab_bench.sh "--steps 20 --warmup 5 --no-extras" noreh default
done
bash scripts/ab_libs.sh "--nlam 5000 200000" noreh default
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000" noreh default
} > gpurun_out/r02_run35.log 2>&1
cat gpurun_out/r02_run35.log
