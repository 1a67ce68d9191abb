"""The two structs that cross the C ABI (include/sphmw.h) have ONE layout: what a C compiler gives
them (tests/c/test_capi.c, plain C11, built against the header and the library), what the ctypes
binding declares (_capi.py) and what the Julia layer declares (julia/SmoothedParticlesB200.jl).
Also: the header is valid C (not only C++) and a plain-C program links against the library."""
import ctypes as C
import json
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CDIR = ROOT / "tests" / "c"
PKG = ROOT / "sph_mountain_waves_b200"


def build_c_program() -> Path:
    from sph_mountain_waves_b200 import _capi
    assert _capi.LIB_PATH.exists(), "libsphmw.so is not built"
    out = CDIR / "test_capi"
    src = CDIR / "test_capi.c"
    deps = [src, ROOT / "include" / "sphmw.h", _capi.LIB_PATH]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(src),
                        "-o", str(out), f"-L{PKG}", "-lsphmw", "-lm", "-lpthread", f"-Wl,-rpath,{PKG}"], check=True)
    return out


@pytest.fixture(scope="module")
def layout():
    exe = build_c_program()
    r = subprocess.run([str(exe), "layout"], capture_output=True, text=True, check=True)
    return json.loads(r.stdout)


def test_ctypes_structs_match_the_c_layout(layout):
    from sph_mountain_waves_b200 import _capi
    assert C.sizeof(_capi.Config) == layout["sizeof_config"]
    for name, off in layout["config"].items():
        assert getattr(_capi.Config, name).offset == off, name
    assert C.sizeof(_capi.LatticeSetup) == layout["sizeof_lattice_setup"]
    for name, off in layout["lattice_setup"].items():
        assert getattr(_capi.LatticeSetup, name).offset == off, name


def test_julia_struct_matches_the_c_layout(layout):
    """SphmwConfig in the Julia layer: same fields, same order, same sizes (Julia lays an isbits
    struct out like C)"""
    src = (PKG / "julia" / "SmoothedParticlesB200.jl").read_text()
    m = re.search(r"struct SphmwConfig[^\n]*\n(.*?)\nend", src, re.S)
    assert m
    sizes = {"NTuple{3,Cdouble}": (24, 8), "Cdouble": (8, 8), "Int64": (8, 8), "Int32": (4, 4)}
    off, fields = 0, {}
    for line in m.group(1).splitlines():
        line = line.split("#")[0].strip()
        if not line:
            continue
        name, typ = [t.strip() for t in line.split("::")]
        size, align = sizes[typ]
        off = (off + align - 1) // align * align
        fields[name] = off
        off += size
    assert fields == layout["config"]
    assert (off + 7) // 8 * 8 == layout["sizeof_config"]
    assert f"{layout['sizeof_config']} bytes" in src


def test_halo_record_size(layout):
    from sph_mountain_waves_b200 import _capi
    assert _capi.lib().sphmw_halo_record_doubles() == layout["halo_record_doubles"] == 13
