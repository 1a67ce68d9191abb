#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki]
    m = re.match(r"(?:void )?(k_\w+)(<[^(]*>)?\(", name)
    short = (m.group(1) + (m.group(2) or "")) if m else name.split("(")[0]
    short = re.sub(r"\(int\)", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0] / 1e6:.3f} |")
print(f"| total | | {tot / 1e6:.3f} | | |")
