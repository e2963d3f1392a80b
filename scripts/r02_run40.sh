#!/bin/bash
# round 2, GPU call 40: one CTA per level after the sweep (post_level_kernel) — suite, time stamps, A/B against the two-stage kernel
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_stamps.so python scripts/stamp_probe.py C2 2>&1 | grep -E "post:|sweep B: (first CTA|last CTA|records)"
for pass in 1 2; do
bash scripts/ab_bench.sh "--steps 20 --warmup 5 --no-extras" oldpost default
done
bash scripts/ab_libs.sh "--nlam 5000 200000" oldpost default
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000" oldpost default
} > gpurun_out/r02_run40.log 2>&1
cat gpurun_out/r02_run40.log
