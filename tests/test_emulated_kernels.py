"""The pair kernels of libsphmw, compiled for the HOST and run one thread at a time (tests/emu/),
against the oracle — no GPU needed.

tests/emu/emu_pairs.cpp includes the very headers nvcc compiles (csrc/pair_list.cuh,
wcsph_ops.cuh, kernels_sph.cuh, cell_gather.cuh, sphmw_internal.h; csrc/grid_setup.cpp) behind a stand-in
<cuda_runtime.h>, builds the cell-sorted layout, and runs the two fused passes of verlet_step!
(wcsph_perturbed_witch.jl:316-331) four ways: the cell walk, the recorded/replayed pair list with
the integer and with the FP64 pre-test, and the packed-record variant.  It exits non-zero unless
all four agree bit for bit; this test then compares the result with the oracle's operator
sequence.  With -ffp-contract=off and the same libm on both sides the strict closures are expected
to match the oracle exactly — neighbour order, accepted set and arithmetic.

This checks the kernels' LOGIC on the CPU; parity of the compiled device code is the job of the
`-m gpu` tests.  Nothing here is a product path.
"""
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from util import load_oracle, n_mismatch, rel_err

ROOT = Path(__file__).resolve().parent.parent
EMU = ROOT / "tests" / "emu"
CSRC = ROOT / "sph_mountain_waves_b200" / "csrc"
PARAMS = ["dt", "g", "c", "gamma", "alpha", "beta", "eps", "eta", "rho0", "R_mass", "R_gas", "T_bg", "rho_floor",
          "P_floor", "z_t", "z_b", "gamma_r", "fluid"]
POST = ["wcsph.reset_density", "wcsph.compute_density", "wcsph.finalize_density", "wcsph.update_smoothing",
        "wcsph.compute_pressure"]


@pytest.fixture(scope="session")
def emu_binary():
    out = EMU / "build" / "emu_pairs"
    out.parent.mkdir(exist_ok=True)
    deps = [EMU / "emu_pairs.cpp", EMU / "cuda_runtime.h", CSRC / "pair_list.cuh", CSRC / "wcsph_ops.cuh",
            CSRC / "kernels_sph.cuh", CSRC / "cell_gather.cuh", CSRC / "sphmw_internal.h", CSRC / "grid_setup.cpp",
            CSRC / "pair_tile.cuh", CSRC / "tile_map.cuh", CSRC / "q_access.cuh"]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wno-attributes", "-DSPHMW_EMU",
                        f"-I{EMU}", f"-I{ROOT / 'include'}", f"-I{CSRC}", str(EMU / "emu_pairs.cpp"),
                        str(CSRC / "grid_setup.cpp"), "-o", str(out)], check=True)
    return out


def run_emulation(binary, tmp_path, case, fields, fast, stride, cx_shift=-1, nsteps=0):
    n = len(fields["m"])
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as fp:
        fp.write(struct.pack("<6i", int(fast), int(stride), int(cx_shift), len(PARAMS), int(nsteps), 0))
        fp.write(struct.pack("<q", n))
        fp.write(struct.pack("<7d", *case.box_min, *case.box_max, case.h))
        for name in PARAMS:
            fp.write(name.encode().ljust(16, b"\0"))
            fp.write(struct.pack("<d", float(case.params.get(name, 0.0))))
        for key in ("x", "v"):
            for k in range(3):
                fp.write(np.ascontiguousarray(fields[key][:, k], dtype="<f8").tobytes())
        for key in ("m", "h", "rho", "rho_p", "type"):
            fp.write(np.ascontiguousarray(fields[key], dtype="<f8").tobytes())
    r = subprocess.run([str(binary), str(inp), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = outp.read_bytes()
    meta = struct.unpack("<6q", raw[:48])
    arr = np.frombuffer(raw[48:], dtype="<f8").reshape(-1, n)
    names = ["rho", "rho_bg", "rho_p", "h", "P_bg", "P_p", "P"]
    out = {k: arr[i] for i, k in enumerate(names)}
    out["v"] = arr[7:10].T
    if nsteps:   # whole steps: only rho, h, v and x are meaningful
        out = {"rho": arr[0], "h": arr[3], "v": arr[7:10].T, "x": arr[10:13].T}
    m = re.search(r"tiled=(\d+)/(\d+)", r.stdout)
    return dict(n=meta[0], dim=meta[1], pairs_density=meta[2], pairs_force=meta[3], overflow=meta[4],
                cx_shift=meta[5], tiled=int(m.group(1)) if m else None, blocks=int(m.group(2)) if m else None), out


def advanced_state(case, warm_steps=1):
    """the oracle after `warm_steps` full steps and the accelerate!/move! of the next one"""
    o = load_oracle(case)
    o.create_cell_list()
    o.step("wcsph", warm_steps)
    o.apply("wcsph.accelerate")
    o.apply("wcsph.move")
    assert len(o) == case.n
    return o, {k: o.field(k) for k in ("x", "v", "m", "h", "rho", "rho_p", "type")}


def finish_step(o):
    o.create_cell_list()
    for op in POST:
        o.apply(op)
    o.apply("wcsph.balance_of_momentum")
    pairs = o.pair_count()
    o.apply("wcsph.accelerate")
    return pairs


CASES = {
    "hill3d": lambda: cases.bell_hill_3d(24, 12, 10, h_m=3000.0, a=8e3, U=20.0),
    "witch2d": lambda: cases.mountain_wave_2d(n_y=24.0, dom_length=80e3, h_m=3000.0, a=10e3, U=20.0),
    # long rows of cells: most blocks of 128 particles sit in one or two chunk rows, so the tiled
    # kernels (csrc/pair_tile.cuh) stage their neighbourhood instead of falling back to the walk
    "long3d": lambda: cases.bell_hill_3d(100, 8, 6, h_m=2000.0, a=8e3, U=20.0),
    "long2d": lambda: cases.mountain_wave_2d(n_y=16.0, dom_length=400e3, h_m=3000.0, a=10e3, U=20.0),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("cx_shift", [-1, 2])
def test_strict_kernels_equal_the_oracle_bit_for_bit(emu_binary, tmp_path, name, cx_shift):
    """cx_shift 2: cells stored in x-chunks of four columns — the physical order must not matter"""
    case = CASES[name]()
    o, fields = advanced_state(case)
    meta, got = run_emulation(emu_binary, tmp_path, case, fields, fast=0, stride=40 if case.dim == 3 else 32,
                              cx_shift=cx_shift)
    pairs = finish_step(o)
    assert meta["n"] == case.n and meta["dim"] == case.dim and meta["overflow"] == 0
    assert meta["pairs_density"] == meta["pairs_force"] == pairs
    if name.startswith("long") and cx_shift < 0:
        assert meta["tiled"] >= 0.6 * meta["blocks"], meta  # the tiled path really ran
    for f in ("rho", "rho_bg", "rho_p", "h", "P_bg", "P_p", "P", "v"):
        assert n_mismatch(got[f], o.field(f)) == 0, f


@pytest.mark.parametrize("name", list(CASES))
def test_fast_kernels_within_tolerance_and_overflow_path(emu_binary, tmp_path, name):
    """FAST_MATH closures (FMA, reciprocals): same pairs, fields within the north star's 1e-10 of the
    oracle under the per-component + element-wise metric of util.rel_err (absolute differences are
    ~1 ulp of |v|, i.e. 2e-13 of the smallest velocity component's scale); a stride of 8 makes every
    particle overflow its list and walk the cells inside the list kernels"""
    case = CASES[name]()
    o, fields = advanced_state(case)
    pairs = None
    for stride in (40, 8):
        meta, got = run_emulation(emu_binary, tmp_path, case, fields, fast=1, stride=stride)
        if pairs is None:
            pairs = finish_step(o)
        assert meta["pairs_density"] == meta["pairs_force"] == pairs
        assert (meta["overflow"] > 0) == (stride == 8)
        for f in ("rho", "h", "P", "v"):
            assert rel_err(got[f], o.field(f)) <= 1e-10, (stride, f)
            assert rel_err(got[f], o.field(f), floor=1.0) <= 1e-12, (stride, f)   # per component only


def test_disordered_particles(emu_binary, tmp_path):
    """a jittered lattice: uneven cell occupancy, candidates on both sides of the cut-off"""
    case = cases.bell_hill_3d(16, 10, 8, h_m=2000.0, a=8e3, U=20.0)
    rng = np.random.default_rng(4)
    dr = case.info["dr"]
    fluid = case.fields["type"] == 0.0
    case.fields["x"][fluid] += rng.uniform(-0.2 * dr, 0.2 * dr, (int(fluid.sum()), 3))
    o, fields = advanced_state(case, warm_steps=0)
    meta, got = run_emulation(emu_binary, tmp_path, case, fields, fast=0, stride=48)
    pairs = finish_step(o)
    assert meta["pairs_force"] == pairs
    for f in ("rho", "h", "P", "v"):
        assert n_mismatch(got[f], o.field(f)) == 0, f


@pytest.mark.parametrize("name", list(CASES))
def test_whole_steps_equal_the_oracle_bit_for_bit(emu_binary, tmp_path, name):
    """25 fused steps (accelerate!, move!, cell list, density pass, force pass + kick), the pair
    passes cycling through walk / list / list_f64 / records: x, v, rho, h of every particle equal
    the oracle's verlet_step! sequence exactly"""
    case = CASES[name]()
    fields = {k: case.fields[k] for k in ("x", "v", "m", "h", "rho", "rho_p", "type")}
    nsteps = 25
    meta, got = run_emulation(emu_binary, tmp_path, case, fields, fast=0, stride=40, nsteps=nsteps)
    o = load_oracle(case)
    o.create_cell_list()
    o.step("wcsph", nsteps)
    assert len(o) == case.n == meta["n"]
    for f in ("x", "v", "rho", "h"):
        assert n_mismatch(got[f], o.field(f)) == 0, f
