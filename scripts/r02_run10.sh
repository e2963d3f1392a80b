#!/bin/bash
mkdir -p gpurun_out
{
bash scripts/ab_bench.sh "--steps 40 --warmup 5 --no-extras" default g4s3 pref g4s3pref default g4s3 pref g4s3pref
for tag in default g4s3; do
  echo "== size_scan $tag"
  if [ "$tag" = default ]; then lib=""; else lib="frei_b200/_lib/variants/libfrei_b200_$tag.so"; fi
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --nlam 37888 151552 200000 800000 2>&1 | grep -E "^L |rror"
done
} > gpurun_out/r02_run10.log 2>&1
cat gpurun_out/r02_run10.log
