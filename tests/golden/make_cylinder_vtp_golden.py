#!/usr/bin/env python
"""Generates tests/golden/cylinder_vtp.json from the reference's own fixture
/root/reference/sph_jl/examples/init/cylinder.vtp (written by WriteVTK.jl, read back by the
reference with import_particles!, sph_jl/examples/cylinder.jl:87) with an INDEPENDENT pure-Python
parser (regex + zlib), so that libsphmw's .vtp reader can be pinned against a file the reference
itself ships.  Only a numeric summary is committed, not the file.

    python tests/golden/make_cylinder_vtp_golden.py
"""
import json
import re
import zlib
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/sph_jl/examples/init/cylinder.vtp")
OUT = Path(__file__).with_name("cylinder_vtp.json")


def parse(path):
    raw = path.read_bytes()
    head, _, rest = raw.partition(b'<AppendedData encoding="raw">')
    blob = rest[rest.index(b"_") + 1:]
    n = int(re.search(rb'NumberOfPoints="(\d+)"', head).group(1))
    out = {}
    for m in re.finditer(rb'<DataArray type="(\w+)" Name="([^"]+)" NumberOfComponents="(\d+)" format="appended" '
                         rb'offset="(\d+)"/>', head):
        typ, name, nc, off = m.group(1).decode(), m.group(2).decode(), int(m.group(3)), int(m.group(4))
        nb, bs, last = np.frombuffer(blob, dtype="<u8", count=3, offset=off)
        sizes = np.frombuffer(blob, dtype="<u8", count=int(nb), offset=off + 24)
        pos = off + 24 + 8 * int(nb)
        data = b""
        for sz in sizes:
            data += zlib.decompress(blob[pos:pos + int(sz)])
            pos += int(sz)
        a = np.frombuffer(data, dtype={"Float64": "<f8", "Int64": "<i8"}[typ]).astype(np.float64)
        out[name] = (nc, a)
    return n, out


def main():
    n, arrays = parse(SRC)
    summary = {"source": str(SRC), "n_points": n, "arrays": {}}
    for name, (nc, a) in arrays.items():
        if name in ("connectivity", "offsets"):
            continue
        summary["arrays"][name] = {
            "ncomp": nc, "sum": float(a.sum()), "sumsq": float((a * a).sum()), "min": float(a.min()),
            "max": float(a.max()), "first": a[:6].tolist(), "last": a[-6:].tolist(),
        }
    vals, counts = np.unique(arrays["type"][1], return_counts=True)
    summary["type_histogram"] = {str(v): int(c) for v, c in zip(vals, counts)}
    OUT.write_text(json.dumps(summary, indent=1))
    print(OUT, n, list(summary["arrays"]))


if __name__ == "__main__":
    main()
