#!/bin/bash
# A/B of sweep-kernel build knobs on the GPU box (nvcc is in the image): kernel-only timing via bench.py
for flags in "$@"; do
  echo "== $flags"
  FREI_B200_NVCC_EXTRA="$flags" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('value %.3e  step %.3f ms  sweep %.4f ms  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_avg_ms'], d['roofline']['frac']))
    elif 'rror' in ln: print(ln.strip())
"
done
