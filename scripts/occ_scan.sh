#!/bin/bash
# time for exactly one wave of sweep CTAs at 1..4 CTAs/SM (128 threads, 256 wavelengths each)
for cfg in "210000 37888" "100000 75776" "60000 113664" "0 151552"; do
  set -- $cfg
  echo "== smem pad $1 B, n_lam $2"
  FREI_B200_NVCC_EXTRA="-DSWEEP_SMEM_PAD=$1" python scripts/size_scan.py --nlam $2 2>&1 | tail -1
done
