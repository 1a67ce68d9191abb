#!/usr/bin/env python
"""Copies the dam-break validation curves the reference ships (digitised from Violeau's SPH book
and from the experiment of Koshizuka & Oka 1996; sph_jl/examples/reference/dambreak_*.csv, plotted
by sph_jl/examples/collapse_dry.jl:233-247) into tests/golden/dambreak_curves.json so that the
validation can run where /root/reference is not mounted.  X = front position / column width,
H = column height / initial height, time = t*sqrt(2|g|/width) (collapse_dry.jl:224).

    python tests/golden/make_dambreak_golden.py
"""
import csv
import json
from pathlib import Path

SRC = Path("/root/reference/sph_jl/examples/reference")
OUT = Path(__file__).with_name("dambreak_curves.json")


def curve(name):
    rows = [r for r in list(csv.reader(open(SRC / name)))[1:] if len(r) == 2 and r[0].strip()]
    pts = sorted((float(a), float(b)) for a, b in rows)
    return {"time": [p[0] for p in pts], "value": [p[1] for p in pts]}


def main():
    out = {"source": str(SRC), "X_Violeau": curve("dambreak_X_Violeau.csv"),
           "X_Koshizuka": curve("dambreak_X_Koshizuka.csv"), "H_Violeau": curve("dambreak_H_Violeau.csv"),
           "H_Koshizuka": curve("dambreak_H_Koshizuka.csv")}
    OUT.write_text(json.dumps(out, indent=1))
    print(OUT, {k: len(v["time"]) for k, v in out.items() if k != "source"})


if __name__ == "__main__":
    main()
