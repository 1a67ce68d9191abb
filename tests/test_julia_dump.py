"""Pins the oracle (and, with a GPU, the device path) against the REFERENCE ITSELF when a dump
made by oracle/dump_state.jl on a machine with Julia is present in tests/golden/julia_dump/.
Without the dump the tests skip: the build image has no `julia`, which is why
DESIGN.md says "parity unpinned"."""
import json
from pathlib import Path

import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from util import load_oracle, rel_err

DUMP = Path(__file__).parent / "golden" / "julia_dump"
pytestmark = pytest.mark.skipif(not (DUMP / "meta.json").exists(),
                                reason="no Julia dump (run oracle/dump_state.jl where julia exists)")


def load(tag):
    f64 = lambda n: np.fromfile(DUMP / f"{n}_{tag}.f64", dtype="<f8")
    i64 = lambda n: np.fromfile(DUMP / f"{n}_{tag}.i64", dtype="<i8")
    return {"x": f64("x").reshape(-1, 3), "v": f64("v").reshape(-1, 3), "rho": f64("rho"), "h": f64("h"),
            "m": f64("m"), "type": f64("type"), "keys": i64("keys"), "pi": i64("pairs_i"), "pj": i64("pairs_j")}


def build(meta):
    return cases.mountain_wave_2d(n_y=meta["n_y"], dom_length=meta["dom_length"])


def check(sysm, ref, tol):
    assert len(sysm) == len(ref["rho"])
    assert np.array_equal(sysm.cell_keys(), ref["keys"])                       # bit-exact cell assignment
    pi, pj = sysm.pairs()
    assert np.array_equal(pi, ref["pi"]) and np.array_equal(pj, ref["pj"])      # bit-exact neighbour lists
    for f in ("x", "v", "rho", "h"):
        assert rel_err(sysm.field(f), ref[f]) <= tol, f


def test_oracle_against_the_julia_reference():
    meta = json.loads((DUMP / "meta.json").read_text())
    case = build(meta)
    ref0 = load("0")
    # the lattice generator and the Particle constructor (grids.jl, geometry.jl, :103-145)
    assert np.array_equal(case.fields["x"], ref0["x"]) and np.array_equal(case.fields["type"], ref0["type"])
    assert rel_err(case.fields["m"], ref0["m"]) <= 1e-15
    o = load_oracle(case)
    o.create_cell_list()
    check(o, ref0, 0.0)
    o.step("wcsph", 1)
    check(o, load("1"), 1e-10)
    o.step("wcsph", meta["nsteps"] - 1)
    check(o, load(str(meta["nsteps"])), 1e-6)


@pytest.mark.gpu
def test_device_against_the_julia_reference(gpu):
    from util import load_gpu
    meta = json.loads((DUMP / "meta.json").read_text())
    s = load_gpu(build(meta))
    s.create_cell_list()
    check(s, load("0"), 0.0)
    s.step(1)
    check(s, load("1"), 1e-10)
    s.step(meta["nsteps"] - 1)
    check(s, load(str(meta["nsteps"])), 1e-6)
