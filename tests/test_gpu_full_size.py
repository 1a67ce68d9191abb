"""BASELINE.json's full-size configurations on the device.  C2 (2D static atmosphere, 244 k
particles) and C3 (2D Witch of Agnesi, 4.1 M particles) are still small enough for a direct
comparison with the oracle; the 3D family is checked at 9.2 M particles through
size-independent properties (momentum conservation of the pair force, idempotence of the
cell list, determinism, slab-count independence)."""
import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from sph_mountain_waves_b200.slabs import LocalCluster, SlabRun
from util import bits_equal, load_gpu, load_oracle, field_err, rel_err

pytestmark = pytest.mark.gpu
TOL_STEP = 1e-10


def momentum_defect(s):
    """sum_p m_p Dv_p of the pair force is zero: the pairwise terms of
    wcsph_perturbed_witch.jl:272,284 are antisymmetric in (p, q) once multiplied by m_p"""
    m, dv = s.field("m"), s.field("Dv")
    tot = np.abs((m[:, None] * dv).sum(axis=0)).max()
    scale = np.abs(m[:, None] * dv).sum()
    return tot / scale


def prepare_force(s):
    for op in ("wcsph.reset_density", "wcsph.compute_density", "wcsph.finalize_density",
               "wcsph.update_smoothing", "wcsph.compute_pressure", "wcsph.balance_of_momentum"):
        s.apply(op)


def test_c2_static_atmosphere_full_size_vs_oracle(gpu):
    case = cases.static_atmosphere_2d()          # dr = 26 km / 120
    assert case.n == 243831 and int((case.fields["type"] == 0).sum()) == 221687
    o, s = load_oracle(case), load_gpu(case)
    assert o.create_cell_list() == s.create_cell_list() == case.n
    assert o.key_tables()[1] == s.key_tables()[1] == (1034, 75, 1)
    assert np.array_equal(o.cell_keys(), s.cell_keys())
    pio, pjo = o.pairs()
    pis, pjs = s.pairs()
    assert np.array_equal(pio, pis) and np.array_equal(pjo, pjs)
    # the pair relation is symmetric
    fwd = set(zip(pis[:200000].tolist(), pjs[:200000].tolist()))
    allp = set(zip(pis.tolist(), pjs.tolist()))
    assert all((j, i) in allp for i, j in fwd)
    o.step("wcsph", 3)
    s.step(3)
    # at rest v is the residual of pressure gradient against buoyancy (the well-balance test,
    # wcsph_perturbed_witch.jl:253-256): measured in units of g*dt per step, not of itself
    for f in ("rho", "v", "x", "h"):
        assert field_err(case, f, s.field(f), o.field(f), 3) <= 3 * TOL_STEP, f


def test_c3_witch_full_size_vs_oracle(gpu):
    case = cases.witch_2d()                      # dr = 26 km / 510, ~4.1 M particles
    assert 4.0e6 < case.n < 4.3e6
    o, s = load_oracle(case), load_gpu(case)
    assert o.create_cell_list() == s.create_cell_list() == case.n
    assert o.key_tables() == s.key_tables()
    assert np.array_equal(o.cell_keys(), s.cell_keys())
    s.count_pairs(True)
    o.step("wcsph", 1)
    s.step(1)
    assert s.pair_count() == o.pair_count()
    for f in ("rho", "v", "x", "h"):
        assert field_err(case, f, s.field(f), o.field(f)) <= TOL_STEP, f
    # idempotence: a second create_cell_list! changes nothing (wcsph_perturbed_witch.jl:320)
    before = {f: s.field(f) for f in ("x", "v", "rho")}
    keys = s.cell_keys()
    assert s.create_cell_list() == case.n
    assert np.array_equal(keys, s.cell_keys())
    for f, a in before.items():
        assert bits_equal(a, s.field(f)), f
    prepare_force(s)
    assert momentum_defect(s) < 1e-10  # rounding of ~1e8 antisymmetric pair terms; one missing pair gives ~1e-7


def test_3d_9M_properties(gpu):
    case = cases.bell_hill_3d(960, 75, 96, lean=True)
    assert case.n > 9.0e6
    s = load_gpu(case)
    assert s.create_cell_list() == case.n
    s.count_pairs(True)
    prepare_force(s)
    pairs = s.pair_count()
    assert 20 * case.n < pairs < 32 * case.n and pairs % 2 == 0      # symmetric relation
    assert momentum_defect(s) < 1e-10  # rounding of ~1e8 antisymmetric pair terms; one missing pair gives ~1e-7
    # determinism + independence of the number of slabs, bit for bit, at this size
    a = load_gpu(case, flags=1)
    a.create_cell_list()
    a.step(2)
    ref = {f: a.field(f) for f in ("x", "v", "rho")}
    a.close()
    cluster = LocalCluster([SlabRun.from_global_case(case, r, 3, flags=1) for r in range(3)])
    cluster.create_cell_list()
    cluster.step(2)
    _, got = cluster.gather(("x", "v", "rho"))
    for f in ref:
        assert np.array_equal(got[f], ref[f]), f


@pytest.mark.parametrize("dims", [(480, 38, 48), (960, 75, 96)], ids=["1.5M", "9.2M"])
@pytest.mark.parametrize("flags", [0, 1], ids=["strict", "fast"])
def test_bell_hill_3d_steps_vs_oracle_at_bench_sizes(gpu, dims, flags):
    """The benched 3D family (BASELINE config 4/5 workloads bell_hill_3d_1M and _8M) against the
    oracle (OpenMP, all host cores): cell keys and pair counts bit-exact, rho / v / x / h within the
    north star's 1e-10 per step after 1 step and within 20e-10 after 20 steps
    (wcsph_perturbed_witch.jl:309-332), strict and fast arithmetic; in strict mode the first
    density sum (no transcendental feeds it) is bit-identical."""
    import os

    from oracle import oracle as O
    case = cases.bell_hill_3d(*dims, lean=True)
    O.set_threads(os.cpu_count() or 1)
    o, s = load_oracle(case), load_gpu(case, flags=flags)
    assert o.create_cell_list() == s.create_cell_list() == case.n
    assert np.array_equal(o.cell_keys(), s.cell_keys())
    s.count_pairs(True)
    o.step("wcsph", 1)
    s.step(1)
    assert s.pair_count() == o.pair_count()
    if flags == 0:
        assert bits_equal(s.field("rho"), o.field("rho"))
    for f in ("rho", "v", "x", "h"):
        assert field_err(case, f, s.field(f), o.field(f)) <= TOL_STEP, f
    o.step("wcsph", 19)
    s.step(19)
    assert len(o) == len(s) == case.n
    assert s.pair_count() == o.pair_count()
    for f in ("rho", "v", "x", "h"):
        assert field_err(case, f, s.field(f), o.field(f), 20) <= 20 * TOL_STEP, f
    s.close()
