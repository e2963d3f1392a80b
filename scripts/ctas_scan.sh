#!/bin/bash
# Sweep-kernel time against resident CTAs per SM (2, 3, 4) over wavelength counts: the occupancy is
# capped with extra dynamic shared memory (build knob SWEEP_SMEM_PAD).  Result of round 1
# (profiles/r01_ctas_scan.log): 4 CTAs/SM are never slower than 3 and up to 18 % faster.
for pad in 46080 24576 0; do
  python frei_b200/build.py --variant pad$pad -DSWEEP_SMEM_PAD=$pad > /dev/null
done
for pad in 46080 24576 0; do
  echo "== SWEEP_SMEM_PAD $pad"
  FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_pad$pad.so python scripts/size_scan.py \
      --nlam 100000 125000 160000 200000 250000 300000 400000 600000 2>&1 | grep -E "^L |rror"
done
