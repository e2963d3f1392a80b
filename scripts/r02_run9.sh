#!/bin/bash
mkdir -p gpurun_out
{
for tag in wide2g4 recg2 recg3; do
  echo "== size_scan $tag (L=100, S=8)"
  lib="frei_b200/_lib/variants/libfrei_b200_$tag.so"
  FREI_B200_LIB=$lib timeout 300 python scripts/size_scan.py --L 100 --S 8 --nlam 75776 125000 250000 1000000 2>&1 | grep -E "^L |rror"
done
echo "== parity of the S=8 cases with recg3"
FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_recg3.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_memory_safety.py -m gpu -q -k "large_shapes or sweeps_match or ragged or guard" 2>&1 | tail -3
} > gpurun_out/r02_run9.log 2>&1
cat gpurun_out/r02_run9.log
