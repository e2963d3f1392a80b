#!/bin/bash
# round 2, GPU call 1: parity of the mixed-width sweep plan, then A/B of the plan and two knobs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_pytest1.log
tail -3 gpurun_out/r02_pytest1.log
{
for tag in default nomix; do
  echo "== size_scan $tag"
  if [ "$tag" = default ]; then lib=""; else lib="frei_b200/_lib/variants/libfrei_b200_$tag.so"; fi
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --nlam 100000 151552 200000 250000 303104 400000 800000 2>&1 | grep -E "^L |rror"
done
echo "== size_scan default, 32-wide chunks only"
timeout 300 python scripts/size_scan.py --plan 1 --nlam 75776 100000 151552 200000 2>&1 | grep -E "^L |rror"
echo "== C3 / C5 shapes per GPU at 8 GPUs (S=8 L=100 125k; L=200 S=3 250k), default vs nomix"
for tag in default nomix; do
  if [ "$tag" = default ]; then lib=""; else lib="frei_b200/_lib/variants/libfrei_b200_$tag.so"; fi
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --L 100 --S 8 --nlam 125000 250000 2>&1 | grep -E "^L |rror"
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --L 200 --S 3 --nlam 250000 2>&1 | grep -E "^L |rror"
done
} > gpurun_out/r02_scan1.log 2>&1
cat gpurun_out/r02_scan1.log
bash scripts/ab_bench.sh "--steps 20 --warmup 3" default nomix emit3 deg5 > gpurun_out/r02_ab1.log 2>&1
cat gpurun_out/r02_ab1.log
