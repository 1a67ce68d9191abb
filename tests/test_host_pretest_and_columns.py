"""Host-only checks (no GPU) of two pieces of logic the device kernels rely on:

* the integer pre-tests of the pair-list recording pass (csrc/sphmw_internal.h: nl_q6_*, one packed
  add + DP4A on the 6-bit mirror of the default cell order; nl_q10_*, the 10-bit mirror of the
  shared-memory tile variant): they may let false candidates through — the exact FP64 test `r > sys.h` (src/core.jl:104-105) follows —
  but it must never reject a pair the reference accepts;
* the column sets of the overlapped slab step (csrc/pair_ops.cu sphmw_slab_cols_of): edge and
  interior columns partition the local grid, the force sets are their owned parts.

Both are reached through the C ABI of libsphmw.so (the same inline functions the kernels
compile), with plain host pointers.
"""
import ctypes as C

import numpy as np
import pytest

from sph_mountain_waves_b200 import _capi

GHOST = 2


def pretest(xp, xq, h, dim, bits=10):
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    xq = np.ascontiguousarray(xq, dtype=np.float64)
    out = np.empty(len(xp), dtype=np.uint8)
    fn = _capi.lib().sphmw_pretest_pairs_q6 if bits == 6 else _capi.lib().sphmw_pretest_pairs
    rc = fn(_capi.ptr(xp), _capi.ptr(xq), len(xp), float(h), dim, _capi.ptr(out))
    assert rc == 0
    return out


def r2_left_to_right(xp, xq, dim):
    """dist() of the reference: dx*dx + dy*dy + dz*dz, left to right, no FMA (core.jl:8-10)"""
    d = xp - xq
    r2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
    if dim == 3:
        r2 = r2 + d[:, 2] * d[:, 2]
    return r2


def r2_max(h):
    """largest double whose correctly rounded square root is <= h (csrc/api.cu, Grid::r2_max)"""
    t = h * h
    while np.sqrt(t) > h:
        t = np.nextafter(t, 0.0)
    while np.sqrt(np.nextafter(t, np.inf)) <= h:
        t = np.nextafter(t, np.inf)
    return t


@pytest.mark.parametrize("bits", [6, 10])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("h,origin", [(1.0, 0.0), (390.0, -2.0e5), (0.0317, 11.0), (2925.0, 3.9e5), (1e-3, -7.0)])
def test_pretest_never_rejects_an_accepted_pair(dim, h, origin, bits):
    rng = np.random.default_rng(int(abs(origin)) + dim)
    n = 400_000
    xp = np.zeros((n, 3))
    xp[:, :dim] = origin + rng.uniform(-40 * h, 40 * h, (n, dim))
    # partners at distances concentrated on the cut-off: exactly h in random directions, a few
    # ulps either side, and a uniform fill of the ball
    u = rng.normal(size=(n, 3))
    u[:, dim:] = 0.0
    u /= np.linalg.norm(u, axis=1)[:, None]
    scale = np.where(rng.random(n) < 0.5, 1.0 + rng.integers(-4, 5, n) * 2.3e-16, rng.random(n) ** (1.0 / dim))
    xq = xp + u * (h * scale)[:, None]
    # grid-aligned partners too: along the axes, where cell faces and the cut-off coincide
    axis = rng.integers(0, dim, n // 4)
    xq[: n // 4] = xp[: n // 4]
    xq[np.arange(n // 4), axis] += h * np.where(rng.random(n // 4) < 0.5, 1.0, -1.0)
    accepted = ~(r2_left_to_right(xp, xq, dim) > r2_max(h))
    assert accepted.sum() > n // 3
    got = pretest(xp, xq, h, dim, bits)
    # an accepted partner is always in one of the 27 cells, and always passes
    assert not np.any(got[accepted] == 2)
    assert np.all(got[accepted] == 1), f"{int(np.sum(got[accepted] == 0))} accepted pairs rejected by the pre-test"


@pytest.mark.parametrize("bits,unit", [(6, 64.0), (10, 1024.0)])
def test_pretest_is_tight(bits, unit):
    """false candidates are confined to a thin shell: beyond (1 + 4/unit) h everything is rejected
    (unit = quantisation steps per cell: h/64 or h/1024)"""
    rng = np.random.default_rng(5)
    n, h = 300_000, 2.5
    xp = rng.uniform(-100.0, 100.0, (n, 3))
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    r = rng.uniform(1.0, 1.5, n) * h
    xq = xp + u * r[:, None]
    got = pretest(xp, xq, h, 3, bits)
    far = r > (1.0 + 4.0 / unit) * h
    assert not np.any(got[far] == 1)
    shell = (got == 1) & (r > h)
    assert shell.sum() > 0 and np.max(r[shell]) < (1.0 + 4.0 / unit) * h


def column_sets(width, has_left, has_right):
    out = (C.c_int32 * 16)()
    assert _capi.lib().sphmw_slab_column_sets(width, int(has_left), int(has_right), out) == 0
    names = ("edge", "interior", "force_edge", "force_interior")
    sets = {}
    for k, name in enumerate(names):
        a0, a1, b0, b1 = out[4 * k:4 * k + 4]
        sets[name] = set(range(a0, a1 + 1)) | set(range(b0, b1 + 1))
    return sets


@pytest.mark.parametrize("width", [8, 9, 10, 11, 12, 13, 20, 144])
@pytest.mark.parametrize("has_left,has_right", [(True, True), (True, False), (False, True)])
def test_slab_column_sets(width, has_left, has_right):
    s = column_sets(width, has_left, has_right)
    cols = set(range(width))
    owned = set(range(GHOST, width - GHOST))
    # every column is advanced exactly once
    assert s["edge"] | s["interior"] == cols and not (s["edge"] & s["interior"])
    # the force pass covers exactly the owned columns, split the same way
    assert s["force_edge"] == s["edge"] & owned and s["force_interior"] == s["interior"] & owned
    # the columns whose particles can produce a record after a drift of less than one column
    # (ghost columns, the two outermost owned ones and the one next to them) are edge columns
    need = set()
    if has_left:
        need |= set(range(0, min(2 * GHOST + 1, width)))
    if has_right:
        need |= set(range(max(width - 2 * GHOST - 1, 0), width))
    assert need <= s["edge"]
    # and nothing else is (the interior is as large as it can be)
    assert s["edge"] == need


def test_slab_column_sets_reject_too_narrow():
    out = (C.c_int32 * 16)()
    assert _capi.lib().sphmw_slab_column_sets(7, 1, 1, out) < 0


def reference_removal(n, removed):
    """create_cell_list! on a vector of labels (src/core.jl:60-81): the removal cell holds the indices in
    DESCENDING order (add_index!, core.jl:26-41), entry i is overwritten by particles[end+1-i], then the
    vector is cut"""
    particles = list(range(n))
    rem = sorted(removed, reverse=True)
    for i, r in enumerate(rem, start=1):
        particles[r] = particles[n - i]
    return particles[: n - len(rem)]


@pytest.mark.parametrize("seed", range(6))
def test_swap_from_end_removal_moves(seed):
    """the index moves the library derives (host replay over a sparse map, csrc/cell_list.cu) against the
    reference's loop executed literally: also when removed particles sit in the tail that is swapped in"""
    rng = np.random.default_rng(seed)
    for n, k in ((1, 1), (2, 1), (10, 10), (50, 7), (1000, 1), (1000, 333), (5000, 4999)):
        removed = rng.choice(n, size=k, replace=False).astype(np.int64)
        if seed % 2:  # crowd the tail: the swapped-in particles are themselves removed ones
            removed = np.unique(np.concatenate([removed[: k // 2], np.arange(n - (k - k // 2), n)])).astype(np.int64)
            k = len(removed)
        old = np.empty(k, dtype=np.int64)
        new = np.empty(k, dtype=np.int64)
        m = C.c_int64()
        rc = _capi.lib().sphmw_swap_removal_moves(n, _capi.ptr(removed), k, _capi.ptr(old), _capi.ptr(new), C.byref(m))
        assert rc == 0 and 0 <= m.value <= k
        want = reference_removal(n, removed.tolist())
        got = np.arange(n)                       # label at each slot
        alive = np.ones(n, dtype=bool)
        alive[removed] = False
        label_at = {}                            # new slot -> label
        moved_from = set(old[: m.value].tolist())
        for o, w in zip(old[: m.value].tolist(), new[: m.value].tolist()):
            assert alive[o] and w < n - k
            label_at[w] = o
        for slot in range(n - k):
            if slot in label_at:
                assert want[slot] == label_at[slot]
            else:
                assert alive[slot] and slot not in moved_from and want[slot] == slot
