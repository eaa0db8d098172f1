#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output per source line:
instructions executed, stall samples and the dominant stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [N]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file, cur_fn, hdr = None, None, None
agg = {}
first_fn = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        first_fn = first_fn or cur_fn
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    extra = len(r) - len(hdr)                      # unquoted commas inside the source text
    if extra > 0:
        r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
    d = dict(zip(hdr, r))
    try:
        inst = float(d.get("Instructions Executed") or 0)
        samp = float(d.get("# Samples") or 0)
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, dict(src=r[1].strip(), inst=0.0, samp=0.0, stalls={}))
    a["inst"] += inst
    a["samp"] += samp
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            try:
                a["stalls"][k] = a["stalls"].get(k, 0.0) + float(v)
            except ValueError:
                pass
tot_i = sum(a["inst"] for a in agg.values()) or 1
tot_s = sum(a["samp"] for a in agg.values()) or 1
print(f"kernel: {first_fn}\ntotal inst {tot_i:.3g}  total samples {tot_s:.0f}")
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s}  top stalls | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{k[6:]}={v:.0f}" for k, v in st if v > 0)
    print(f"{key[0] + ':' + str(key[1]):28s} {100 * a['inst'] / tot_i:6.2f} {100 * a['samp'] / tot_s:6.2f}  {sts} | {a['src'][:90]}")
