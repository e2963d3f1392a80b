"""Instruction mix of the layer loop of a sweep kernel in a built library (cuobjdump -sass).
usage: python scripts/sass_loop_mix.py LIB [mangled-kernel-substring]
Finds the largest backward branch of the kernel and counts opcodes between target and branch
(both E paths of the warp vote are inside, so the counts are an upper bound of the executed path)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else 'sweep_kernelIdLi3ELi0ELi2ELb0'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, rows = None, []
for ln in txt.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
best = (0, 0, 0)
for a, ins in rows:
    m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)', ins)
    if m:
        t = int(m.group(1), 16)
        if t < a and a - t > best[0]:
            best = (a - t, t, a)
_, lo, hi = best
c = collections.Counter()
for a, ins in rows:
    if lo <= a <= hi:
        ins = re.sub(r'^@!?U?P\d+\s+', '', ins)
        c[ins.split()[0].split('.')[0]] += 1
fp64 = c['DFMA'] + c['DMUL'] + c['DADD']
print(f'{lib}: loop {lo:#x}..{hi:#x}: {sum(c.values())} instr, fp64 {fp64} '
      f'(DFMA {c["DFMA"]} DMUL {c["DMUL"]} DADD {c["DADD"]}), MUFU {c["MUFU"]}, LDS {c["LDS"]}, '
      f'LDGSTS {c["LDGSTS"]}, SHFL {c["SHFL"]}, IMAD {c["IMAD"]}, FSEL {c["FSEL"]}')
