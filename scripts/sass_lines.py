#!/usr/bin/env python
"""Per-source-line SASS statistics of one kernel from `nvdisasm --print-line-info` output:
instruction count, spill (STL/LDL) count — where do the registers run out."""
import collections
import re
import sys

path, kern = sys.argv[1], sys.argv[2]
cur = None
on = False
cnt, tot, ops = collections.Counter(), collections.Counter(), collections.Counter()
for l in open(path):
    if l.startswith("//---") and ".text." in l:
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.search(r'^\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m:
        tot[cur] += 1
        op = m.group(2).split('.')[0]
        ops[op] += 1
        if op in ('STL', 'LDL'):
            cnt[cur] += 1
print("total instructions", sum(tot.values()), "spill ops", sum(cnt.values()))
print("top opcodes:", ops.most_common(24))
print("lines with spills:")
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:20]:
    print("  ", k, "spill", v, "of", tot[k])
print("largest lines:")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:25]:
    print("  ", k, v)
