#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv` launch list per kernel.
usage: python profiles/summarize.py launches.csv"""
import collections
import csv
import re
import sys

TIME = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, ui, vi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launches = collections.OrderedDict()
for r in data:
    if len(r) > vi:
        launches.setdefault((r[0], r[ki]), {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
agg = collections.OrderedDict()
for (_, name), m in launches.items():
    mm = re.match(r"(?:void )?(k_\w+)(<[^(]*>)?\(", name)
    short = (mm.group(1) + (mm.group(2) or "")) if mm else name.split("(")[0]
    short = re.sub(r"\(int\)|\(bool\)", "", short)
    a = agg.setdefault(short, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    t, u = m["gpu__time_duration.sum"]
    a[1] += t * TIME.get(u, 1e-6)
    for k, slot in (("dram__bytes_read.sum", 2), ("dram__bytes_write.sum", 3)):
        if k in m:
            a[slot] += m[k][0] * BYTES.get(m[k][1], 1.0)
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | share | avg ms | DRAM read GB / launch | DRAM write GB / launch |")
print("|---|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.3f} | "
          f"{v[2] / v[0] / 1e9:.3f} | {v[3] / v[0] / 1e9:.3f} |")
print(f"| total | | {tot:.3f} | | | | |")
