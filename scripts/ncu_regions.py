#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` metrics over named source regions.
usage: python scripts/ncu_regions.py src.csv name:file:lo-hi[,file:lo-hi] ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; cur = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    extra = len(r) - len(hdr)
    if extra > 0: r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
    d = dict(zip(hdr, r))
    try: inst = float(d['Instructions Executed'] or 0); s = float(d['# Samples'] or 0)
    except ValueError: continue
    k = (cur, int(r[0])); a = agg.get(k, (0, 0)); agg[k] = (a[0] + inst, a[1] + s)
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
seen = set()
for spec in sys.argv[2:]:
    name, rest = spec.split(':', 1)
    i = s = 0
    for part in rest.split(','):
        f, rng = part.split(':'); lo, hi = map(int, rng.split('-'))
        for k, v in agg.items():
            if k[0].startswith(f) and lo <= k[1] <= hi and k not in seen:
                i += v[0]; s += v[1]; seen.add(k)
    print(f"{name:26s} inst {100*i/tot:5.1f}% ({i/1e6:7.1f}M)  samples {100*s/ts:5.1f}%")
i = sum(v[0] for k, v in agg.items() if k not in seen); s = sum(v[1] for k, v in agg.items() if k not in seen)
print(f"{'(other)':26s} inst {100*i/tot:5.1f}% ({i/1e6:7.1f}M)  samples {100*s/ts:5.1f}%   total {tot/1e6:.1f}M")
