#!/bin/bash
# Sweep-kernel throughput vs resident CTAs per SM: extra dynamic shared memory (SWEEP_SMEM_PAD) caps
# the occupancy, the persistent grid follows it.  Builds the variants in place (nvcc is in the image),
# so run it where a GPU is: scripts/occ_scan.sh
for pad in 150000 80000 40000 0; do
  python frei_b200/build.py --variant occ$pad -DSWEEP_MINB=4 -DSWEEP_SMEM_PAD=$pad > /dev/null
done
scripts/ab_libs.sh "--nlam 200000 800000" occ150000 occ80000 occ40000 occ0
