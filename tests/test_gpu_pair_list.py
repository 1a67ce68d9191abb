"""The pair list (include/sphmw.h SPHMW_FLAG_*PAIR_LIST*, csrc/pair_list.cuh): recording the
candidates of _apply_binary! (src/core.jl:94-112) on the first binary pass of a cell list and
replaying them on the later ones must not change a single bit — same accepted set (the exact
FP64 test `r > sys.h` is repeated on every entry), same traversal order."""
import os

import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from test_gpu_edge_cases import tiny_case
from util import bits_equal, load_gpu, load_oracle, n_mismatch

pytestmark = pytest.mark.gpu

FAST_MATH, NO_LIST, EAGER, NO_PRETEST = 1, 4, 8, 16
FIELDS = ["x", "v", "h", "rho", "rho_p", "rho_bg", "P", "P_p", "P_bg", "T", "theta", "type"]


def small_2d():
    return cases.mountain_wave_2d(n_y=24.0, dom_length=80e3, h_m=3000.0, a=10e3, U=20.0)


def small_3d():
    return cases.bell_hill_3d(24, 12, 10, h_m=3000.0, a=8e3, U=20.0)


@pytest.mark.parametrize("make", [small_2d, small_3d])
@pytest.mark.parametrize("arith", [0, FAST_MATH])
def test_fused_step_same_bits_with_and_without_list(gpu, make, arith):
    """list: the recording pass pre-tests candidates with integers on the 10-bit cell-relative
    mirror; list_f64 (NO_PRETEST): with the exact FP64 test only"""
    case = make()
    runs = {}
    for name, flags in (("walk", NO_LIST), ("list", 0), ("list_f64", NO_PRETEST)):
        s = load_gpu(case, flags=flags | arith)
        s.create_cell_list()
        s.count_pairs(True)
        s.step(6)
        runs[name] = ({f: s.field(f) for f in FIELDS}, s.pair_count(), s.pair_list_info())
    assert runs["walk"][2]["builds"] == 0
    for name in ("list", "list_f64"):
        info = runs[name][2]
        assert info["builds"] == 6 and info["overflow"] == 0, info
        assert runs[name][1] == runs["walk"][1] > 0
        for f in FIELDS:
            assert bits_equal(runs[name][0][f], runs["walk"][0][f]), (name, f)


@pytest.mark.parametrize("make", [small_2d, small_3d])
def test_overflowing_particles_walk_the_cells(gpu, make, monkeypatch):
    """a stride far too small for the lattice: every particle overflows and takes the
    original loop inside the recording and the replaying kernels"""
    case = make()
    ref = load_gpu(case, flags=NO_LIST)
    ref.create_cell_list()
    ref.step(3)
    for stride in ("4", "12"):
        monkeypatch.setenv("SPHMW_PAIR_LIST_STRIDE", stride)
        s = load_gpu(case)
        s.create_cell_list()
        s.step(3)
        info = s.pair_list_info()
        assert info["stride"] == int(stride) and (info["overflow"] > 0 or stride == "12")
        for f in FIELDS:
            assert bits_equal(s.field(f), ref.field(f)), (stride, f)


def test_queue_rows_grow_when_particles_overflow(gpu, monkeypatch):
    """the recording pass's shared-memory queue starts far too short here: particles overflow and walk
    the cells (same bits), the library sees the overflow counter grow and adds rows, four per cell
    list, until the lists fit again"""
    monkeypatch.setenv("SPHMW_PAIR_QUEUE_ROWS", "16")
    case = small_3d()
    ref = load_gpu(case, flags=NO_LIST)
    ref.create_cell_list()
    s = load_gpu(case)
    s.create_cell_list()
    s.step(2)
    first = s.pair_list_info()["overflow"]
    assert first > 0
    s.step(10)
    settled = s.pair_list_info()["overflow"]
    s.step(4)
    assert s.pair_list_info()["overflow"] == settled   # no particle overflows any more
    ref.step(16)
    for f in FIELDS:
        assert bits_equal(s.field(f), ref.field(f)), f


@pytest.mark.parametrize("flags", [EAGER, EAGER | NO_PRETEST])
def test_operator_by_operator_with_eager_list(gpu, flags):
    """op-by-op through apply!: the density pass records, the force pass replays; sums stay
    bit-identical to the oracle where no transcendental is involved"""
    case = small_3d()
    o, s = load_oracle(case), load_gpu(case, flags=flags)
    o.create_cell_list()
    s.create_cell_list()
    s.count_pairs(True)
    for op in ("wcsph.reset_density", "wcsph.compute_density"):
        o.apply(op)
        s.apply(op)
    assert o.pair_count() == s.pair_count()
    assert n_mismatch(s.field("rho"), o.field("rho")) == 0
    for op in ("wcsph.finalize_density", "wcsph.update_smoothing", "wcsph.compute_pressure"):
        o.apply(op)
    for f in ("rho_p", "rho_bg", "h", "P", "P_p", "P_bg"):
        s.set_field(f, o.field(f))
    o.apply("wcsph.balance_of_momentum")
    s.apply("wcsph.balance_of_momentum")
    assert o.pair_count() == s.pair_count()
    assert n_mismatch(s.field("Dv"), o.field("Dv")) == 0
    assert s.pair_list_info()["builds"] == 1


@pytest.mark.parametrize("variant", ["hopkins", "hopkins_total"])
def test_three_pass_schemes_pick_the_list_up_by_themselves(gpu, variant):
    """schemes with three binary passes per cell list (hopkins_perturbed_witch.jl:325-349): the
    operator-by-operator sequence (hopkins_total) records a list from the second step on (the
    first cell list shows how many passes use it), the fused step (hopkins) records in every step"""
    case = cases.hopkins_2d(variant)
    a, b = load_gpu(case, flags=NO_LIST), load_gpu(case)
    for s in (a, b):
        s.create_cell_list()
        s.step(5, variant)
    assert a.pair_list_info()["builds"] == 0 and b.pair_list_info()["builds"] == (5 if variant == "hopkins" else 4)
    for f in ("x", "v", "rho", "P", "h"):
        assert bits_equal(a.field(f), b.field(f)), f


def test_crowded_cell_and_cutoff_boundary(gpu):
    """700 particles in one cell (far beyond any stride) and pairs exactly at r == h"""
    rng = np.random.default_rng(2)
    n = 700
    x = np.zeros((n, 3))
    x[:, :2] = rng.uniform(0.05, 0.95, (n, 2))
    x[::7, :2] += 1.0
    case = tiny_case(x, m=rng.uniform(0.5, 1.5, n))
    o, s = load_oracle(case), load_gpu(case, flags=EAGER)
    o.create_cell_list()
    s.create_cell_list()
    for op in ("wcsph.reset_density", "wcsph.compute_density"):
        o.apply(op)
        s.apply(op)
    assert n_mismatch(s.field("rho"), o.field("rho")) == 0
    assert s.pair_list_info()["overflow"] > 0
    # r == h accepted, one ulp beyond rejected — also through the integer pre-test, far from the origin
    off = 2000.0
    x = [[off, 0.0, 0.0], [off + 1.0, 0.0, 0.0], [off, np.nextafter(1.0, 2.0), 0.0],
         [off - 0.6, 0.8, 0.0], [off + 0.3, -0.4, 0.0]]
    case = tiny_case(x, box=((-3.0, -3.0, 0.0), (off + 3.0, 3.0, 0.0)), m=np.arange(1.0, 6.0))
    o, s = load_oracle(case), load_gpu(case, capacity=16, flags=EAGER)
    o.create_cell_list()
    s.create_cell_list()
    s.count_pairs(True)
    for _ in range(2):  # second round replays the list
        for op in ("wcsph.reset_density", "wcsph.compute_density"):
            o.apply(op)
            s.apply(op)
        assert o.pair_count() == s.pair_count()
        assert n_mismatch(s.field("rho"), o.field("rho")) == 0


def test_random_cloud_matches_oracle_pairs_through_the_list(gpu):
    """disordered particles (uneven cell occupancy): accepted pairs counted by the recording
    and the replaying kernels equal the oracle's pair dump"""
    rng = np.random.default_rng(11)
    n = 4000
    x = np.zeros((n, 3))
    x[:, :2] = rng.uniform(-2.9, 2.9, (n, 2))
    case = tiny_case(x, h=0.25, m=rng.uniform(0.5, 1.5, n))
    o, s = load_oracle(case), load_gpu(case, flags=EAGER)
    o.create_cell_list()
    s.create_cell_list()
    s.count_pairs(True)
    pio, _ = o.pairs()
    for _ in range(2):
        for op in ("wcsph.reset_density", "wcsph.compute_density"):
            o.apply(op)
            s.apply(op)
        assert s.pair_count() == len(pio)
        assert n_mismatch(s.field("rho"), o.field("rho")) == 0


@pytest.mark.parametrize("flags", [EAGER, EAGER | NO_PRETEST])
def test_wrapped_cells_of_a_narrow_grid(gpu, flags):
    """quirk 5 (no per-axis bounds check, core.jl:98): on a grid two cells wide the linear key
    arithmetic reaches wrapped cells whose particles CAN be within h; the recording pass must
    find them too (they take the exact test instead of the cell-relative integer one)"""
    rng = np.random.default_rng(9)
    n = 60
    x = np.zeros((n, 3))
    x[:, 0] = rng.uniform(0.0, 1.9, n)
    x[:, 1] = rng.uniform(0.0, 4.9, n)
    case = tiny_case(x, box=((0.0, 0.0, 0.0), (1.95, 4.95, 0.0)), m=rng.uniform(0.5, 1.5, n))
    o, s = load_oracle(case), load_gpu(case, capacity=128, flags=flags)
    o.create_cell_list()
    s.create_cell_list()
    s.count_pairs(True)
    pio, _ = o.pairs()
    for _ in range(2):
        for op in ("wcsph.reset_density", "wcsph.compute_density"):
            o.apply(op)
            s.apply(op)
        assert s.pair_count() == len(pio)
        assert n_mismatch(s.field("rho"), o.field("rho")) == 0


def test_cutoff_boundary_in_3d_across_cell_faces(gpu):
    """pairs exactly at r == h (accepted) and one ulp beyond (rejected) whose partners sit in
    different cells along every axis, far from the origin and at negative coordinates"""
    base = np.array([-37.25, 11.5, -3.75])
    dirs = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.6, 0.8, 0.0], [0.0, -0.6, 0.8], [-0.8, 0.0, 0.6],
                     [1 / 3, 2 / 3, 2 / 3], [-2 / 3, 1 / 3, -2 / 3]])
    pts = [base]
    for dvec in dirs:
        pts.append(base + dvec)                                   # r == h up to rounding
        pts.append(base + dvec * np.nextafter(1.0, 2.0) * (1 + 1e-12))  # just outside
        pts.append(base + dvec * (1 - 1e-12))                     # just inside
    x = np.array(pts)
    n = len(x)
    case = tiny_case(x, box=((-40.0, -3.0, -8.0), (3.0, 15.0, 3.0)), m=np.arange(1.0, n + 1.0))
    assert case.dim == 3
    o, s = load_oracle(case), load_gpu(case, capacity=64, flags=EAGER)
    assert o.create_cell_list() == s.create_cell_list() == n
    s.count_pairs(True)
    pio, pjo = o.pairs()
    assert len(pio) > 2 * len(dirs)
    for _ in range(2):
        for op in ("wcsph.reset_density", "wcsph.compute_density"):
            o.apply(op)
            s.apply(op)
        assert s.pair_count() == len(pio)
        assert n_mismatch(s.field("rho"), o.field("rho")) == 0


# ---------------------------------------------------------------------------------------
# Variants of the fused pair passes.  Default: packed neighbour records (three 32-byte records,
# 256-bit loads); NO_RECORDS: eleven SoA gathers; TILES: neighbourhood of each block staged in
# shared memory (csrc/pair_tile.cuh).  All must give the same bits.
# ---------------------------------------------------------------------------------------
NO_RECORDS = 128
PACKED = 32    # forces the records in 2D, where SoA gathers are the default
TILES = 64


def long_3d():
    """rows of ~60 cells: most blocks of 128 particles sit in one or two chunk rows, so the tiled
    kernels stage their neighbourhood instead of falling back to the cell walk"""
    return cases.bell_hill_3d(100, 8, 6, h_m=2000.0, a=8e3, U=20.0)


def long_2d():
    return cases.mountain_wave_2d(n_y=16.0, dom_length=400e3, h_m=3000.0, a=10e3, U=20.0)


@pytest.mark.parametrize("make", [small_2d, small_3d, long_2d, long_3d])
@pytest.mark.parametrize("arith", [0, FAST_MATH])
@pytest.mark.parametrize("variant", [NO_RECORDS, PACKED, TILES])
def test_pair_pass_variants_same_bits(gpu, make, arith, variant):
    """records (default) vs SoA gathers vs shared-memory tiles: the neighbour data are bit copies
    of the SoA fields and the visiting order is the same, so nothing may change"""
    case = make()
    a, b = load_gpu(case, flags=arith), load_gpu(case, flags=arith | variant)
    for s in (a, b):
        s.create_cell_list()
        s.count_pairs(True)
        s.step(6)
    assert a.pair_count() == b.pair_count() > 0
    for f in FIELDS:
        assert bits_equal(a.field(f), b.field(f)), f
    if variant == TILES and make in (long_2d, long_3d):
        t = b.pair_list_info()["tiles"]
        assert t["staged"] >= 0.6 * t["blocks"], t  # the tiled path really ran


@pytest.mark.parametrize("variant", [0, TILES])
def test_pair_pass_variants_on_slabs_with_the_overlapped_schedule(gpu, variant):
    from sph_mountain_waves_b200.slabs import LocalCluster, SlabRun
    case = cases.bell_hill_3d(48, 10, 8, h_m=3000.0, a=8e3, U=20.0)
    whole = load_gpu(case, flags=FAST_MATH | NO_RECORDS)
    whole.create_cell_list()
    whole.step(12)
    cluster = LocalCluster([SlabRun.from_global_case(case, r, 3, flags=FAST_MATH | variant) for r in range(3)])
    cluster.create_cell_list()
    cluster.step(12)
    _, got = cluster.gather(("x", "v", "rho", "h"))
    for f, arr in got.items():
        assert np.array_equal(arr, whole.field(f)), f


@pytest.mark.parametrize("variant", [0, TILES])
def test_pair_pass_variants_with_overflowing_lists(gpu, monkeypatch, variant):
    monkeypatch.setenv("SPHMW_PAIR_LIST_STRIDE", "12")
    case = long_3d() if variant == TILES else small_3d()
    a, b = load_gpu(case, flags=NO_LIST), load_gpu(case, flags=variant)
    for s in (a, b):
        s.create_cell_list()
        s.step(3)
    assert b.pair_list_info()["overflow"] > 0
    for f in FIELDS:
        assert bits_equal(a.field(f), b.field(f)), f
