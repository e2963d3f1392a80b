#!/bin/bash
# A/B of prebuilt experiment libraries on the whole bench step (kernel + post + launch gaps).
# usage: scripts/ab_bench.sh "<bench args>" TAG...
args="$1"; shift
for tag in "$@"; do
  echo "== $tag"
  if [ "$tag" = default ]; then lib=""; else lib="frei_b200/_lib/variants/libfrei_b200_$tag.so"; fi
  FREI_B200_LIB=$lib python bench.py --no-cpu-baseline $args 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('value %.4e  step %.4f ms  sweep %.4f ms  e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['e2e']['value']))
"
done
