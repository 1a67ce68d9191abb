"""Edge cases of the cell list and the pair traversal (what the reference's semantics imply for
empty, tiny, degenerate and boundary inputs), device against oracle."""
import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from util import bits_equal, load_gpu, load_oracle, n_mismatch

pytestmark = pytest.mark.gpu


def tiny_case(x, h=1.0, box=((-3.0, -3.0, 0.0), (3.0, 3.0, 0.0)), **extra):
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3)
    n = len(x)
    base = cases.mountain_wave_2d(n_y=8.0, dom_length=20e3)
    f = {"x": x, "v": np.zeros((n, 3)), "m": np.ones(n), "h": np.full(n, h), "rho": np.ones(n),
         "rho_p": np.zeros(n), "type": np.zeros(n)}
    f.update(extra)
    return cases.Case("tiny", "wcsph", 2 if box[0][2] == box[1][2] == 0.0 else 3, box[0], box[1], h,
                      dict(base.params), f)


def test_empty_system(gpu):
    case = tiny_case(np.zeros((0, 3)))
    s = load_gpu(case, capacity=16)
    assert s.create_cell_list() == 0
    s.step(2)
    assert len(s) == 0 and s.field("x").shape == (0, 3)
    assert s.reduce("count") == 0.0


def test_single_particle_has_no_neighbours(gpu):
    """no self term (apply! default self=false): rho = 0 after the density pass, and the
    isolated FLUID particle gets a non-finite velocity and is removed at the next cell list,
    exactly as in the reference (SURVEY §5 'failure detection' row)"""
    case = tiny_case([[0.2, 0.3, 0.0]])
    o, s = load_oracle(case), load_gpu(case, capacity=16)
    assert o.create_cell_list() == s.create_cell_list() == 1
    for sysm in (o, s):
        sysm.apply("wcsph.reset_density")
        sysm.apply("wcsph.compute_density")
    assert s.field("rho")[0] == o.field("rho")[0] == 0.0
    o.step("wcsph", 2)
    s.step(2)
    assert len(o) == len(s) == 0


def test_r_equal_h_is_accepted_and_box_is_closed(gpu):
    """quirk 9: r == h passes `r > sys.h`; quirk 6: the bounding box is a closed interval"""
    x = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0],            # exactly h apart
         [0.0, np.nextafter(1.0, 2.0), 0.0],           # one ulp beyond h from particle 0
         [3.0, 3.0, 0.0], [-3.0, -3.0, 0.0],           # exactly on the box corners: inside
         [np.nextafter(3.0, 4.0), 0.0, 0.0]]           # one ulp outside: removed
    case = tiny_case(x)
    o, s = load_oracle(case), load_gpu(case, capacity=16)
    assert o.create_cell_list() == s.create_cell_list() == 5
    pio, pjo = o.pairs()
    pis, pjs = s.pairs()
    assert np.array_equal(pio, pis) and np.array_equal(pjo, pjs)
    assert (0, 1) in set(zip(pis.tolist(), pjs.tolist()))
    assert (0, 2) not in set(zip(pis.tolist(), pjs.tolist()))
    assert bits_equal(o.field("x"), s.field("x"))


def test_many_particles_in_one_cell(gpu):
    """a crowded cell: the in-cell ordering and the pair loops must cope with long runs"""
    rng = np.random.default_rng(2)
    n = 700
    x = np.zeros((n, 3))
    x[:, :2] = rng.uniform(0.05, 0.95, (n, 2))          # all in the cell [0,1) x [0,1)
    x[::7, :2] += 1.0                                     # and some in the diagonal neighbour
    case = tiny_case(x, m=rng.uniform(0.5, 1.5, n), v=np.concatenate([rng.normal(size=(n, 2)), np.zeros((n, 1))], axis=1))
    o, s = load_oracle(case), load_gpu(case)
    assert o.create_cell_list() == s.create_cell_list() == n
    assert np.array_equal(o.cell_keys(), s.cell_keys())
    key = int(s.cell_keys()[1])
    assert np.array_equal(o.cell_entries(key), s.cell_entries(key)) and len(s.cell_entries(key)) > 500
    for op in ("wcsph.reset_density", "wcsph.compute_density"):
        o.apply(op)
        s.apply(op)
    assert n_mismatch(s.field("rho"), o.field("rho")) == 0
    pio, pjo = o.pairs()
    pis, pjs = s.pairs()
    assert np.array_equal(pio, pis) and np.array_equal(pjo, pjs) and len(pis) > 100000


def test_narrow_grids_visit_wrapped_cells_like_the_reference(gpu):
    """quirk 5: no per-axis bounds check — on a grid only two cells wide the linear key
    arithmetic wraps into the neighbouring row and the same pair can be visited twice"""
    rng = np.random.default_rng(9)
    n = 60
    x = np.zeros((n, 3))
    x[:, 0] = rng.uniform(0.0, 1.9, n)
    x[:, 1] = rng.uniform(0.0, 4.9, n)
    case = tiny_case(x, box=((0.0, 0.0, 0.0), (1.95, 4.95, 0.0)))
    o, s = load_oracle(case), load_gpu(case, capacity=128)
    assert o.key_tables()[1][0] == 2
    assert o.create_cell_list() == s.create_cell_list() == n
    pio, pjo = o.pairs()
    pis, pjs = s.pairs()
    assert np.array_equal(pio, pis) and np.array_equal(pjo, pjs)
    for op in ("wcsph.reset_density", "wcsph.compute_density"):
        o.apply(op)
        s.apply(op)
    assert n_mismatch(s.field("rho"), o.field("rho")) == 0
