"""Every operator of the device menu, compiled for the HOST (tests/emu/emu_ops.cpp) and applied by
name, against the oracle — no GPU needed.

The harness includes csrc/ops_menu.cuh: the functors nvcc compiles and SPHMW_OPERATOR_MENU, the
very list pair_ops.cu builds its dispatch table from, so a name reaches the same functor as in
libsphmw.  Binary operators go through the cell walk, the recording kernel (integer or FP64
pre-test) or the replaying kernel of csrc/pair_list.cuh, cycling with the cell-list generation.
With -ffp-contract=off and the same libm, every field must come out with the oracle's bits:
neighbour order, accepted set and arithmetic of the drivers' closures
(src/current/*.jl, sph_jl/examples/collapse_dry.jl, sph_jl/tests/test_collision_2d.jl,
src/utils/new_packing.jl, src/legacy/isothermal_flow_witch.jl).

This checks the LOGIC of the device code on the CPU; `-m gpu` tests check the compiled code.
"""
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from util import load_oracle, n_mismatch

ROOT = Path(__file__).resolve().parent.parent
EMU = ROOT / "tests" / "emu"
CSRC = ROOT / "sph_mountain_waves_b200" / "csrc"

# enum Slot of csrc/sphmw_internal.h
_enum = re.search(r"enum Slot : int \{(.*?)NSLOT", (CSRC / "sphmw_internal.h").read_text(), re.S).group(1)
_enum = re.sub(r"//[^\n]*", "", _enum)
SLOTS = {name.strip().split("=")[0].strip(): i for i, name in enumerate(x for x in _enum.split(",") if x.strip())}
NSLOT = len(SLOTS)
FIELD_SLOT = {"h": ("S_H", 1), "x": ("S_X0", 3), "m": ("S_M", 1), "v": ("S_V0", 3), "Dv": ("S_DV0", 3),
              "rho_bg": ("S_RHO_BG", 1), "rho_p": ("S_RHO_P", 1), "rho": ("S_RHO", 1), "P_bg": ("S_P_BG", 1),
              "P_p": ("S_P_P", 1), "P": ("S_P", 1), "theta_bg": ("S_TH_BG", 1), "theta_p": ("S_TH_P", 1),
              "theta": ("S_TH", 1), "T_bg": ("S_T_BG", 1), "T_p": ("S_T_P", 1), "T": ("S_T", 1),
              "type": ("S_TYPE", 1), "A": ("S_A", 1), "A_bg": ("S_A_BG", 1), "Drho": ("S_DRHO", 1),
              "rho0": ("S_RHO0", 1), "S": ("S_ENT", 1), "s": ("S_ENT_D", 1)}

# the operator sequences of one verlet_step! per scheme (csrc/pair_ops.cu sphmw_step_scheme)
WCSPH = ["wcsph.accelerate", "wcsph.move", "create_cell_list", "wcsph.reset_density", "wcsph.compute_density",
         "wcsph.finalize_density", "wcsph.update_smoothing", "create_cell_list", "wcsph.compute_pressure",
         "wcsph.find_temperature", "wcsph.find_pot_temp", "wcsph.balance_of_momentum", "wcsph.accelerate"]
HOPKINS = ["wcsph.accelerate", "wcsph.move", "create_cell_list", "wcsph.reset_density", "wcsph.compute_density",
           "wcsph.finalize_density", "wcsph.update_smoothing", "hopkins.reset_pressure", "hopkins.compute_pressure",
           "hopkins.finalize_pressure", "wcsph.find_temperature", "wcsph.find_pot_temp",
           "wcsph.balance_of_momentum", "wcsph.accelerate"]
SEQUENCES = {
    "wcsph": WCSPH,
    "hopkins": HOPKINS,
    "hopkins_full": HOPKINS[:-2] + ["hopkins_full.balance_of_momentum", "wcsph.accelerate"],
    "hopkins_total": ["hopkins_total.accelerate", "hopkins_total.move", "create_cell_list",
                      "hopkins_total.reset_density", "wcsph.compute_density", "wcsph.update_smoothing",
                      "hopkins_total.reset_pressure", "hopkins.compute_pressure", "hopkins_total.finalize_pressure",
                      "hopkins_total.find_temperature", "hopkins_total.find_pot_temp",
                      "hopkins_total.balance_of_momentum", "hopkins_total.accelerate"],
    "dambreak": ["dambreak.accelerate", "dambreak.move", "create_cell_list", "dambreak.balance_of_mass",
                 "dambreak.find_pressure", "dambreak.move", "create_cell_list", "dambreak.internal_force",
                 "dambreak.accelerate"],
    "collision": ["collision.accelerate", "collision.move", "create_cell_list", "collision.reset_rho",
                  "+collision.find_rho", "collision.find_pressure", "collision.reset_a", "collision.internal_force",
                  "collision.accelerate"],
    "packing": ["packing.reset_rho", "packing.accumulate_rho", "packing.balance_of_momentum", "packing.accelerate",
                "packing.move", "create_cell_list"],
    "flow": ["flow.accelerate", "flow.move", "create_cell_list", "flow.balance_of_mass", "flow.find_pressure",
             "flow.find_pot_temp", "flow.internal_force", "flow.accelerate"],
    # src/legacy/adiabatic_flow_witch.jl:231-243 (add_new_particles! is host logic around a device kernel:
    # tests/test_gpu_more_schemes.py)
    "aflow": ["flow.accelerate", "aflow.move", "create_cell_list", "+aflow.find_density", "aflow.find_s",
              "aflow.find_pressure", "aflow.entropy_production", "flow.internal_force", "flow.accelerate"],
}


@pytest.fixture(scope="session")
def emu_ops_binary():
    out = EMU / "build" / "emu_ops"
    out.parent.mkdir(exist_ok=True)
    deps = [EMU / "emu_ops.cpp", EMU / "cuda_runtime.h", CSRC / "ops_menu.cuh", CSRC / "wcsph_ops.cuh",
            CSRC / "pair_list.cuh", CSRC / "kernels_sph.cuh", CSRC / "cell_gather.cuh", CSRC / "sphmw_internal.h",
            CSRC / "grid_setup.cpp"]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wno-attributes", "-DSPHMW_EMU",
                        f"-I{EMU}", f"-I{ROOT / 'include'}", f"-I{CSRC}", str(EMU / "emu_ops.cpp"),
                        str(CSRC / "grid_setup.cpp"), "-o", str(out)], check=True)
    return out


def emulate(binary, tmp_path, case, fields, ops, stride=48, cx_shift=-1):
    n = len(next(iter(fields.values())))
    arrays = []
    for name, a in fields.items():
        slot, ncomp = FIELD_SLOT[name]
        a = np.asarray(a, dtype=np.float64)
        for k in range(ncomp if case.dim == 3 or ncomp == 1 else 2):
            arrays.append((SLOTS[slot] + k, a[:, k] if ncomp == 3 else a))
    inp, outp = tmp_path / "ops_in.bin", tmp_path / "ops_out.bin"
    with open(inp, "wb") as fp:
        fp.write(struct.pack("<4i", stride, cx_shift, len(case.params), len(arrays)))
        fp.write(struct.pack("<q", n))
        fp.write(struct.pack("<7d", *case.box_min, *case.box_max, case.h))
        for name, value in case.params.items():
            fp.write(name.encode().ljust(16, b"\0"))
            fp.write(struct.pack("<d", float(value)))
        for slot, a in arrays:
            fp.write(struct.pack("<i", slot))
            fp.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
        fp.write(struct.pack("<i", len(ops)))
        for op in ops:
            self_ = op.startswith("+")
            fp.write(op.lstrip("+").encode().ljust(48, b"\0"))
            fp.write(struct.pack("<i", int(self_)))
    r = subprocess.run([str(binary), str(inp), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = outp.read_bytes()
    meta = struct.unpack("<4q", raw[:32])
    arr = np.frombuffer(raw[32:], dtype="<f8").reshape(NSLOT, n)
    out = {}
    for name, (slot, ncomp) in FIELD_SLOT.items():
        s = SLOTS[slot]
        out[name] = arr[s:s + 3].T if ncomp == 3 else arr[s]
    return dict(n=meta[0], dim=meta[1], pairs=meta[2], overflow=meta[3], log=r.stdout), out


def run_oracle(case, ops, prologue=()):
    o = load_oracle(case)
    o.create_cell_list()
    for op in list(prologue) + list(ops):
        if op == "create_cell_list":
            assert o.create_cell_list() == case.n
        else:
            o.apply(op.lstrip("+"), op.startswith("+"))
    return o


def compare_all_fields(case, got, o):
    for name in FIELD_SLOT:
        want = o.field(name)
        have = got[name]
        if case.dim == 2 and have.ndim == 2:
            have = np.column_stack([have[:, :2], np.zeros(len(have))])
        assert n_mismatch(have, want) == 0, name


CASES = {
    "wcsph": lambda: cases.bell_hill_3d(16, 10, 8, h_m=2000.0, a=8e3, U=20.0),
    "hopkins": lambda: cases.hopkins_2d("hopkins"),
    "hopkins_full": lambda: cases.hopkins_2d("hopkins_full"),
    "hopkins_total": lambda: cases.hopkins_2d("hopkins_total"),
    "dambreak": lambda: cases.collapse_dry(dr=4e-2),
    "collision": cases.collision_2d,
    "flow": lambda: cases.flow_2d(n_y=20.0, dom_length=30e3, h_m=4e3, a=4e3, U_max=40.0),
    "aflow": lambda: cases.aflow_2d(n_y=20.0, dom_length=30e3, h_m=4e3, a=4e3, U_max=40.0),
}
# what the drivers run once before the time loop
PROLOGUE = {"dambreak": ["dambreak.internal_force"],
            # what closes make_system() of the adiabatic driver (:121-126)
            "aflow": ["aflow.find_density", "aflow.find_pressure", "aflow.find_pot_temp", "aflow.find_s",
                      "flow.internal_force"],
            "collision": ["+collision.find_rho0", "+collision.find_rho", "collision.find_pressure",
                          "collision.internal_force"]}


@pytest.mark.parametrize("scheme", list(CASES))
def test_scheme_steps_equal_the_oracle_bit_for_bit(emu_ops_binary, tmp_path, scheme):
    """four steps of each driver's verlet_step!, operator by operator and by name"""
    case = CASES[scheme]()
    nsteps = 4
    ops = ["create_cell_list"] + PROLOGUE.get(scheme, []) + SEQUENCES[scheme] * nsteps
    meta, got = emulate(emu_ops_binary, tmp_path, case, case.fields, ops)
    o = run_oracle(case, SEQUENCES[scheme] * nsteps, prologue=PROLOGUE.get(scheme, []))
    assert meta["n"] == len(o) == case.n and meta["dim"] == case.dim
    assert meta["pairs"] == o.pair_count()
    compare_all_fields(case, got, o)


def test_packing_operators(emu_ops_binary, tmp_path):
    """one pseudo-step of packing! (src/utils/new_packing.jl:96-108)"""
    case = cases.mountain_wave_2d(n_y=20.0, dom_length=60e3)
    k = case.params
    for name, val in (("dt_pack", 1.0 * k["dt"]), ("c_pack", 2.0 * k["c"]), ("zeta_pack", 1.0 * k["c"] / k["dt"])):
        case.params[name] = val
    ops = ["create_cell_list"] + SEQUENCES["packing"] * 3
    meta, got = emulate(emu_ops_binary, tmp_path, case, case.fields, ops)
    o = run_oracle(case, SEQUENCES["packing"] * 3)
    compare_all_fields(case, got, o)


def test_every_menu_entry_is_exercised():
    """the sequences above reach every operator libsphmw lists (sphmw_op_list), except the inflow
    spawning of the flow driver, which is host logic around a device kernel"""
    import ctypes as C
    from sph_mountain_waves_b200 import _capi
    buf = C.create_string_buffer(16384)
    _capi.lib().sphmw_op_list(buf, 16384)
    menu = {line.split()[0] for line in buf.value.decode().splitlines() if line.strip()}
    used = {op.lstrip("+") for seq in SEQUENCES.values() for op in seq} | \
           {op.lstrip("+") for seq in PROLOGUE.values() for op in seq}
    used.discard("create_cell_list")
    assert used <= menu
    assert menu - used <= {"flow.set_density"}, sorted(menu - used)
