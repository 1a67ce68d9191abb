"""x-slab decomposition on the device: any number of ranks must give the
SAME bits as the whole-domain run — neighbours are visited in (cell, global index) order."""
import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from sph_mountain_waves_b200.slabs import LocalCluster, SlabRun
from util import load_gpu, load_oracle, field_err, rel_err

pytestmark = pytest.mark.gpu


def wide_2d():
    return cases.mountain_wave_2d(n_y=16.0, dom_length=120e3, h_m=3000.0, a=10e3, U=20.0)


def wide_3d():
    return cases.bell_hill_3d(48, 10, 8, h_m=3000.0, a=8e3, U=20.0)


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("make,world,flags", [(wide_2d, 2, 0), (wide_2d, 3, 0), (wide_3d, 2, 0), (wide_3d, 4, 0),
                                              (wide_3d, 3, 1), (wide_2d, 2, 2), (wide_3d, 2, 4)])
def test_slabs_bitwise_equal_to_whole_domain(gpu, make, world, flags, overlap):
    """flags: 0 strict, 1 FAST_MATH, 2 CELL_PAIRS, 4 NO_PAIR_LIST — each path is rank-count
    independent; overlap: the schedule that packs the edge columns first and sends their records
    while the interior is in the force pass (sphmw_step_phase 2/3) gives the same bits"""
    if overlap and flags == 2:
        pytest.skip("the cell-pairs kernel has no overlapped schedule")
    case = make()
    whole = load_gpu(case, flags=flags)
    whole.create_cell_list()
    cluster = LocalCluster([SlabRun.from_global_case(case, r, world, flags=flags) for r in range(world)])
    cluster.create_cell_list()
    assert sum(r.n_owned for r in cluster.runs) == case.n
    nsteps = 25  # U = 20 m/s: particles do cross slab faces within these steps
    whole.step(nsteps)
    cluster.step(12, overlap=overlap)   # two calls: the state between them is a plain step boundary
    cluster.step(nsteps - 12, overlap=overlap)
    gidx, got = cluster.gather(("x", "v", "rho", "h", "rho_p", "m", "type"))
    assert np.array_equal(gidx, np.arange(case.n))
    for f, a in got.items():
        assert np.array_equal(a, whole.field(f)), f
    assert sum(r.n_owned for r in cluster.runs) == case.n


def test_particles_migrate_between_slabs(gpu):
    """a strong wind moves particles across the slab face; ownership must follow"""
    case = cases.mountain_wave_2d(n_y=16.0, dom_length=120e3, U=250.0)
    world = 2
    cluster = LocalCluster([SlabRun.from_global_case(case, r, world) for r in range(world)])
    cluster.create_cell_list()
    before = [r.n_owned for r in cluster.runs]
    own0 = set(cluster.runs[0].owned_fields(("m",))[0].tolist())
    whole = load_gpu(case)
    whole.create_cell_list()
    whole.step(60)
    cluster.step(60)   # overlapped schedule
    after = [r.n_owned for r in cluster.runs]
    own1 = set(cluster.runs[0].owned_fields(("m",))[0].tolist())
    assert sum(after) == sum(before) == case.n
    assert own0 != own1, "no particle changed owner: the test does not exercise migration"
    gidx, got = cluster.gather(("x", "v", "rho"))
    for f, a in got.items():
        assert np.array_equal(a, whole.field(f)), f


def test_slab_run_against_oracle(gpu):
    case = wide_3d()
    o = load_oracle(case)
    o.create_cell_list()
    o.step("wcsph", 5)
    cluster = LocalCluster([SlabRun.from_global_case(case, r, 3) for r in range(3)])
    cluster.create_cell_list()
    cluster.step(5)
    _, got = cluster.gather(("x", "v", "rho", "h"))
    for f, a in got.items():
        assert field_err(case, f, a, o.field(f)) <= 1e-10, f


def test_overlapped_step_refuses_fast_particles(gpu):
    """the overlapped schedule packs the edge columns before the interior has moved: a particle
    crossing more than one cell column in a step must make it fail loudly, not lose a ghost"""
    from sph_mountain_waves_b200._capi import SphmwError
    case = cases.mountain_wave_2d(n_y=16.0, dom_length=120e3, U=20.0)
    h = case.h
    dt = case.params["dt"]
    v = case.fields["v"]
    x = case.fields["x"]
    # a fluid particle seven columns right of the slab face moving three columns per step towards
    # it: the plain first step takes it to the interior side of rank 1's edge columns, the first
    # overlapped step from there into the columns whose ghost records were already packed
    from sph_mountain_waves_b200.slabs import plan_slab
    plan = plan_slab(case.box_min, case.box_max, h, 1, 2)
    face = (plan.phase0 + plan.lo) * h
    cand = np.nonzero((case.fields["type"] == 0.0) & (x[:, 0] > face + 7.2 * h) & (x[:, 0] < face + 7.8 * h)
                      & (x[:, 1] > 5e3) & (x[:, 1] < 8e3))[0]
    assert len(cand) > 0
    v[cand[0], 0] = -3.0 * h / dt
    cluster = LocalCluster([SlabRun.from_global_case(case, r, 2) for r in range(2)])
    cluster.create_cell_list()
    with pytest.raises(SphmwError, match="more than one cell column"):
        cluster.step(4)


@pytest.mark.parametrize("variant,world,U", [("hopkins", 2, 20.0), ("hopkins", 3, 250.0), ("hopkins_full", 3, 20.0)])
def test_hopkins_slabs_bitwise_equal_to_whole_domain(gpu, variant, world, U):
    """the pressure-entropy drivers on slabs (hopkins_perturbed_witch.jl:324-349,
    full_hopkins_perturbed_witch.jl:350-374): three pair passes per step; compute_pressure! reads the
    neighbours' NEW smoothing length, so a slab keeps three ghost columns (SPHMW_FLAG_GHOST3) and the
    halo records carry A (and A_bg).  Any number of ranks gives the whole-domain bits; U = 250 m/s makes
    particles change owner on the way."""
    case = cases.hopkins_2d(variant, n_y=16.0, dom_length=120e3, h_m=3000.0, a=10e3, U=U)
    whole = load_gpu(case)
    whole.create_cell_list()
    cluster = LocalCluster([SlabRun.from_global_case(case, r, world) for r in range(world)])
    cluster.create_cell_list()
    assert sum(r.n_owned for r in cluster.runs) == case.n
    own0 = [r.n_owned for r in cluster.runs]
    nsteps = 30
    whole.step(nsteps)
    cluster.step(nsteps)
    names = ("x", "v", "rho", "h", "P", "A", "m", "type") + (("A_bg",) if variant == "hopkins_full" else ())
    gidx, got = cluster.gather(names)
    assert np.array_equal(gidx, np.arange(case.n))
    for f, a in got.items():
        assert np.array_equal(a, whole.field(f)), f
    assert sum(r.n_owned for r in cluster.runs) == case.n
    if U > 100.0:
        assert [r.n_owned for r in cluster.runs] != own0, "no particle changed owner"


def test_hopkins_on_a_two_column_slab_is_refused(gpu):
    """without the third ghost column the pressure sums next to the slab face would be incomplete:
    the step fails loudly instead"""
    from sph_mountain_waves_b200._capi import SphmwError
    case = cases.hopkins_2d("hopkins", n_y=16.0, dom_length=120e3)
    case.scheme = "wcsph"          # sliced like a WCSPH case: two ghost columns
    run = SlabRun.from_global_case(case, 0, 2)
    run.sys.T.scheme = "hopkins"
    cluster = LocalCluster([run])
    with pytest.raises(SphmwError):
        run.backend.pre()
        run.backend.post()
