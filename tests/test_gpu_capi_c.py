"""The library driven from plain C (tests/c/test_capi.c): create -> upload -> create_cell_list ->
step -> download without Python in the loop, compared with the same scenario through the ctypes
binding and the oracle; and, with two GPUs, x-slabs stepped by the unchanged sphmw_step call (halo
transport inside the library) from two host threads, bit-identical to the one-GPU run."""
import json
import subprocess

import numpy as np
import pytest

from test_capi_layout import build_c_program
from util import load_gpu, load_oracle, n_mismatch

pytestmark = pytest.mark.gpu


def c_scenario():
    """the particle cloud of test_capi.c: make_cloud(40, 10, 8), constants of set_params()"""
    from sph_mountain_waves_b200 import cases
    from sph_mountain_waves_b200.schemes import wcsph_perturbed_witch as wpw
    k = wpw.Constants(n_y=12.0)
    dr = k.dr
    nx, ny, nz = 40, 10, 8
    i, j, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i, j, kk = i.ravel(), j.ravel(), kk.ravel()
    x = np.stack([(i + 0.5) * dr, (j + 0.5) * dr, (kk + 0.5) * dr], axis=1)
    v = np.stack([20.0 + 5.0 * ((7 * j + 3 * kk) % 11) / 11.0, 0.5 * ((9 * i) % 7) / 7.0 - 0.25, np.zeros(len(i))], axis=1)
    rho = k.rho0 / (1.0 + x[:, 1] * k.g / (k.R_mass * k.T_bg))
    fields = dict(x=x, v=v, m=rho * dr * dr * dr, h=np.full(len(i), k.h0), rho=rho, rho_p=np.zeros(len(i)),
                  type=np.zeros(len(i)))
    return cases.Case("c_scenario", "wcsph", 3, (0.0, 0.0, 0.0), (nx * dr, ny * dr, nz * dr), k.h0, k.params(),
                      fields, {})


def fnv(a: np.ndarray) -> str:
    s = 0
    for b in np.ascontiguousarray(a, dtype="<f8").view("<u8").ravel().tolist():
        s = (s * 1099511628211 + b) & 0xFFFFFFFFFFFFFFFF
    return f"{s:016x}"


def test_one_gpu_from_plain_c(gpu):
    exe = build_c_program()
    r = subprocess.run([str(exe), "run"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    case = c_scenario()
    assert got["n0"] == got["alive"] == got["n"] == case.n and got["launches"] > 0
    s = load_gpu(case)
    s.create_cell_list()
    s.step(3)
    # component-major, as the C program downloads them
    assert got["x_bits"] == fnv(s.field("x").T) and got["v_bits"] == fnv(s.field("v").T)
    assert got["rho_bits"] == fnv(s.field("rho"))
    o = load_oracle(case)
    o.create_cell_list()
    o.step("wcsph", 3)
    assert abs(got["rho_first"] - o.field("rho")[0]) <= 1e-10 * abs(o.field("rho")[0])
    assert n_mismatch(s.field("rho"), o.field("rho")) < case.n  # (exp feeds rho from step 2 on: not all bits equal)


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_from_plain_c_threads(gpu, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    exe = build_c_program()
    r = subprocess.run([str(exe), "slabs", str(world)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    assert got["bitwise_equal_to_one_gpu"] is True and got["owned"] == got["n"]
