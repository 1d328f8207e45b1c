#!/usr/bin/env python
"""Aggregate the warp-stall samples of an ncu report by CUDA source line.
   ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top [kernel]]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
only = sys.argv[3] if len(sys.argv) > 3 else ""   # optional: restrict to kernels whose name contains this
cur_file, cur_fn, hdr, out = None, "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split('/')[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif r[0] != "" and hdr is not None and len(r) > 10 and r[0].isdigit() and only in cur_fn:
        out.append((cur_file, r))
idx = {n: i for i, n in enumerate(hdr)}
si = hdr.index("# Samples")
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def cell(r, i):
    return r[i] if i < len(r) else ""


agg = {}
for f, r in out:
    key = (f, int(r[0]), r[1].strip()[:100])
    a = agg.setdefault(key, [0, {}])
    a[0] += num(cell(r, si))
    for n in stalls:
        v = num(cell(r, idx[n]))
        if v:
            a[1][n] = a[1].get(n, 0) + v
tot = sum(a[0] for a in agg.values())
print('total samples', tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[1].items(), key=lambda kv: -kv[1])[:3]
    print(f"{k[0]}:{k[1]:4d} {a[0]:7d} {100*a[0]/tot:5.1f}% {st} | {k[2]}")
