#!/bin/bash
# A/B of prebuilt experiment libraries (frei_b200/build.py --variant TAG FLAGS...): sweep-kernel
# timing over a few wavelength counts.  usage: scripts/ab_libs.sh "<size_scan args>" TAG...
args="$1"; shift
for tag in "$@"; do
  echo "== $tag"
  if [ "$tag" = default ]; then python scripts/size_scan.py $args 2>&1 | grep "^L "
  else FREI_B200_LIB=frei_b200/_lib/variants/libfrei_b200_$tag.so python scripts/size_scan.py $args 2>&1 | grep "^L "; fi
done
