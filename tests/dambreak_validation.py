"""Shared by the CPU (oracle) and GPU dam-break validations: run collapse_dry
(sph_jl/examples/collapse_dry.jl:194-230) to t* = 3.2 and compare the surge-front position X(t*)
and the column height H(t*) (get_globals, :171-186) with the curves the reference ships."""
import json
import math
from pathlib import Path

import numpy as np

CURVES = json.loads((Path(__file__).parent / "golden" / "dambreak_curves.json").read_text())


def run_dambreak(system, case, step, every=50, t_star_end=3.2):
    """`system` has .field(name); `step(n)` advances n steps.  Returns (t*, X, H) samples."""
    p = case.params
    dt, h = p["dt"], p["kh"]
    scale = math.sqrt(-2 * p["gy"])               # collapse_dry.jl:224
    nsteps = int(t_star_end / scale / dt)
    ts, Xs, Hs = [], [], []
    for k in range(0, nsteps, every):
        step(every)
        x, typ = system.field("x"), system.field("type")
        fluid = typ == 0.0
        Xs.append(x[fluid, 0].max() / 1.0)                       # water_column_width = 1
        sel = fluid & (x[:, 0] < 2.0) & (x[:, 0] > h)
        Hs.append(x[sel, 1].max() / 2.0)                         # water_column_height = 2
        ts.append((k + every) * dt * scale)
    return np.array(ts), np.array(Xs), np.array(Hs)


def deviation(name, ts, ys, t_max=None):
    c = CURVES[name]
    t, v = np.array(c["time"]), np.array(c["value"])
    m = t <= (ts[-1] if t_max is None else min(ts[-1], t_max))
    sim = np.interp(t[m], ts, ys)
    return float(np.max(np.abs(sim / v[m] - 1.0))), int(m.sum())
