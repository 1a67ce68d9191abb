"""x-slab decomposition on the device: any number of ranks must give the
SAME bits as the whole-domain run — neighbours are visited in (cell, global index) order."""
import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from sph_mountain_waves_b200.slabs import LocalCluster, SlabRun
from util import load_gpu, load_oracle, rel_err

pytestmark = pytest.mark.gpu


def wide_2d():
    return cases.mountain_wave_2d(n_y=16.0, dom_length=120e3, h_m=3000.0, a=10e3, U=20.0)


def wide_3d():
    return cases.bell_hill_3d(48, 10, 8, h_m=3000.0, a=8e3, U=20.0)


@pytest.mark.parametrize("make,world,flags", [(wide_2d, 2, 0), (wide_2d, 3, 0), (wide_3d, 2, 0), (wide_3d, 4, 0),
                                              (wide_3d, 3, 1), (wide_2d, 2, 2)])
def test_slabs_bitwise_equal_to_whole_domain(gpu, make, world, flags):
    """flags: 0 strict, 1 FAST_MATH, 2 CELL_PAIRS — each path is rank-count independent"""
    case = make()
    whole = load_gpu(case, flags=flags)
    whole.create_cell_list()
    cluster = LocalCluster([SlabRun.from_global_case(case, r, world, flags=flags) for r in range(world)])
    cluster.create_cell_list()
    assert sum(r.n_owned for r in cluster.runs) == case.n
    nsteps = 25  # U = 20 m/s: particles do cross slab faces within these steps
    whole.step(nsteps)
    cluster.step(nsteps)
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
    cluster.step(60)
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
        assert rel_err(a, o.field(f)) <= 1e-10, f
