#!/bin/bash
# round 2, GPU call 32: next-cell prefetch of the table rows (second staging buffer) — suite, then A/B against one buffer
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for pass in 1 2; do
bash scripts/ab_libs.sh "--nlam 37888 100000 151552 200000 303104 800000" rb1 default
done
bash scripts/ab_libs.sh "--L 100 --S 8 --nlam 125000 1000000" rb1 default
bash scripts/ab_libs.sh "--L 200 --S 3 --nlam 250000" rb1 default
bash scripts/ab_libs.sh "--table-dtype 32 --nlam 200000 800000" rb1 default
bash scripts/ab_libs.sh "--table-dtype 32 --L 100 --S 8 --nlam 125000 1000000" rb1 default
} > gpurun_out/r02_run32.log 2>&1
cat gpurun_out/r02_run32.log
