"""The C-ABI library loads and exports every symbol include/sphmw.h declares
(no compute call: this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "sphmw.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphmw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sph_mountain_waves_b200 import _capi
    assert _capi.LIB_PATH.exists(), "libsphmw.so is not built (python -m sph_mountain_waves_b200.build)"
    lib = ctypes.CDLL(str(_capi.LIB_PATH))
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sphmw.h but not exported"
    # and the Python binding covers the same surface
    assert sorted(_capi.exported_symbols()) == names


def test_version_and_op_menu_without_gpu():
    from sph_mountain_waves_b200 import _capi, op_menu
    assert b"sm_100a" in _capi.lib().sphmw_version()
    menu = op_menu()
    for op, kind in (("wcsph.compute_density", "binary"), ("wcsph.balance_of_momentum", "binary"),
                     ("wcsph.accelerate", "unary"), ("wcsph.move", "unary")):
        assert menu[op] == kind


def test_no_cpu_fallback_without_device():
    """without a CUDA device the product fails loudly instead of computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sph_mountain_waves_b200 import SphmwError, cases
    s = cases.to_system(cases.mountain_wave_2d(n_y=8.0, dom_length=20e3))
    with pytest.raises(SphmwError) as e:
        s.create_cell_list()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = ROOT / "sph_mountain_waves_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        if "build/" in str(p):
            continue
        txt = p.read_text(errors="ignore")
        assert "import oracle" not in txt and "from oracle" not in txt and "sph_oracle" not in txt, p
