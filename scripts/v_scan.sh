#!/bin/bash
for flags in "-DSWEEP_V=4 -DSWEEP_MINB=2" "-DSWEEP_V=4 -DSWEEP_MINB=3" "-DSWEEP_V=4 -DSWEEP_MINB=2 -DSWEEP_THREADS=64" "-DSWEEP_V=2 -DSWEEP_MINB=4" "-DSWEEP_V=1 -DSWEEP_MINB=6"; do
  echo "== $flags"
  FREI_B200_NVCC_EXTRA="$flags" python scripts/size_scan.py --nlam 200000 800000 2>&1 | tail -2
done
