#!/bin/bash
mkdir -p gpurun_out
{
echo "== parity with deferred reduction"
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_defer1.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for tag in default defer1; do
  echo "== size_scan $tag"
  if [ "$tag" = default ]; then lib=""; else lib="frei_b200/_lib/variants/libfrei_b200_$tag.so"; fi
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --nlam 37888 151552 200000 303104 800000 2>&1 | grep -E "^L |rror"
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --L 100 --S 8 --nlam 125000 2>&1 | grep -E "^L |rror"
done
bash scripts/ab_bench.sh "--steps 20 --warmup 3" default defer1 defer1emit3
echo "== C4 capped atmospheres"
timeout 600 python scripts/c4_capped.py 2>&1 | tail -12
} > gpurun_out/r02_run3.log 2>&1
cat gpurun_out/r02_run3.log
