"""The oracle's driver closures against an INDEPENDENT numpy restatement of the same Julia source
(brute-force O(N^2) neighbour search, vectorised formulas, no cell list).  The summation order
differs, so agreement is to rounding (1e-12), which is what a transcription slip in either
restatement would break by many orders of magnitude."""
import numpy as np
import pytest

from sph_mountain_waves_b200 import cases
from util import load_oracle, rel_err


def neighbours(x, h):
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt((d * d).sum(-1))
    mask = (r <= h) & ~np.eye(len(x), dtype=bool)        # core.jl:105
    return d, r, mask


def wendland2(h, r):                                      # kernels.jl:108-115
    x = r / h
    return np.where(x > 1.0, 0.0, 7 / np.pi * (1 - x) ** 4 * (1 + 4 * x) / h ** 2)


def rDwendland2(h, r):                                    # kernels.jl:140-147
    x = r / h
    return np.where(x > 1.0, 0.0, -140 / np.pi * (1 - x) ** 3 / h ** 4)


def wendland3(h, r):                                      # kernels.jl:156-163
    x = r / h
    return np.where(x > 1.0, 0.0, 21 / (2 * np.pi) * (1 - x) ** 4 * (1 + 4 * x) / h ** 3)


def rDwendland3(h, r):                                    # kernels.jl:188-195
    x = r / h
    return np.where(x > 1.0, 0.0, -210 / np.pi * (1 - x) ** 3 / h ** 5)


def _accelerate(f, p, Dv):                               # :298-303, :245-256
    fluid = f["type"] == p["fluid"]
    ey = np.array([0.0, 1.0, 0.0])
    buoy = -p["g"] * ey[None, :] * (f["rho_p"] / f["rho"])[:, None]
    sn = np.sin(np.pi / 2 * (1 - (p["z_t"] - p["z_b"]) / p["z_b"]))
    damp = np.where((f["x"][:, 1] >= p["z_t"] - p["z_b"])[:, None], -p["gamma_r"] * sn ** 2 * ey[None, :], 0.0)
    f["v"] = np.where(fluid[:, None], f["v"] + 0.5 * p["dt"] * (Dv + buoy + damp), f["v"])


def numpy_step_kick_drift(f, p):
    """accelerate! + move!, wcsph_perturbed_witch.jl:311-312"""
    _accelerate(f, p, f["Dv"])
    fluid = f["type"] == p["fluid"]
    f["x"] = np.where(fluid[:, None], f["x"] + p["dt"] * f["v"], f["x"])            # move! :292-296


def neighbour_pairs(x, h):
    """ordered pairs (i, j), i != j, with |x_i - x_j| <= h (core.jl:104-105), found with a k-d tree
    (scipy) instead of the cell list; the distance test itself is redone in plain FP64"""
    from scipy.spatial import cKDTree
    tree = cKDTree(x)
    und = tree.query_pairs(h * (1 + 1e-9), output_type="ndarray")
    i = np.concatenate([und[:, 0], und[:, 1]])
    j = np.concatenate([und[:, 1], und[:, 0]])
    d = x[i] - x[j]
    r = np.sqrt((d * d).sum(-1))
    keep = r <= h
    return i[keep], j[keep], d[keep], r[keep]


def numpy_step_sums(f, p, h_cut, dim3):
    """:316-331 — density, smoothing length, pressure, pair force, second accelerate!"""
    W, rDW = (wendland3, rDwendland3) if dim3 else (wendland2, rDwendland2)
    n = len(f["m"])
    i, j, d, r = neighbour_pairs(f["x"], h_cut)
    # compute_density! :226-228 (no self term), finalize_density! :230-233, update_smoothing! :235-238
    f["rho"] = np.bincount(i, weights=f["m"][j] * W(f["h"][i], r), minlength=n)
    rho_bg = p["rho0"] * np.exp(-f["x"][:, 1] * p["g"] / (p["R_mass"] * p["T_bg"]))   # :177-179
    f["rho_p"] = f["rho"] - rho_bg
    rfl = np.maximum(f["rho"], p["rho_floor"])
    f["h"] = p["eta"] * (np.cbrt(f["m"] / rfl) if dim3 else np.sqrt(f["m"] / rfl))
    # compute_pressure! :195-199
    P_p = p["c"] ** 2 * f["rho_p"]
    P = p["R_mass"] * p["T_bg"] * rho_bg + P_p
    # balance_of_momentum! :261-286
    dot = (d * (f["v"][i] - f["v"][j])).sum(-1)
    h_ij = 0.5 * (f["h"][i] + f["h"][j])
    ker = rDW(h_ij, r)
    pr = P_p / rfl ** 2
    fc = -f["m"][j] * (pr[i] + pr[j]) * ker
    cs = np.sqrt(p["gamma"] * P / rfl)
    mu = h_ij * dot / (r * r + p["eps"] * h_ij * h_ij)
    pi_ij = (-p["alpha"] * 0.5 * (cs[i] + cs[j]) * mu + p["beta"] * mu * mu) / (0.5 * (rfl[i] + rfl[j]))
    fv = np.where(dot < 0.0, -f["m"][j] * pi_ij * ker, 0.0)
    Dv = np.stack([np.bincount(i, weights=(fc + fv) * d[:, k], minlength=n) for k in range(3)], axis=1)
    _accelerate(f, p, Dv)
    f["Dv"] = np.zeros_like(f["v"])                       # accelerate! zeroes Dv (:302)
    f["P"], f["P_p"], f["Dv_last"] = P, P_p, Dv
    return len(i)


def numpy_wcsph_step(case):
    """verlet_step!, src/current/wcsph_perturbed_witch.jl:309-332, on plain arrays"""
    f = {k: v.copy() for k, v in case.fields.items()}
    numpy_step_kick_drift(f, case.params)
    npairs = numpy_step_sums(f, case.params, case.h, case.dim == 3)
    return f, npairs


@pytest.mark.parametrize("make", [
    lambda: cases.mountain_wave_2d(n_y=12.0, dom_length=30e3, h_m=3000.0, a=6e3, U=25.0),
    lambda: cases.bell_hill_3d(10, 8, 6, h_m=3000.0, a=4e3, U=25.0),
])
def test_wcsph_step_against_numpy_restatement(make):
    case = make()
    # perturb the lattice so that the artificial-viscosity branch and varying h are exercised
    rng = np.random.default_rng(11)
    fluid = case.fields["type"] == 0.0
    case.fields["v"] = case.fields["v"] + fluid[:, None] * rng.normal(scale=8.0, size=(case.n, 3)) * \
        np.array([1.0, 1.0, 1.0 if case.dim == 3 else 0.0])
    case.fields["h"] = case.fields["h"] * rng.uniform(0.9, 1.3, case.n)
    o = load_oracle(case)
    o.create_cell_list()
    o.step("wcsph", 1)
    ref, npairs = numpy_wcsph_step(case)
    assert o.pair_count() == npairs > 0
    for name in ("x", "v", "rho", "rho_p", "h", "P", "P_p"):
        assert rel_err(o.field(name), ref[name]) < 1e-12, name


def test_hopkins_pressure_and_total_force_against_numpy():
    """hopkins_perturbed_witch.jl:205-214 and hopkins_total_witch.jl:233-264"""
    case = cases.hopkins_2d("hopkins_total", n_y=12.0, dom_length=30e3)
    rng = np.random.default_rng(5)
    case.fields["v"] = case.fields["v"] + rng.normal(scale=5.0, size=(case.n, 3)) * np.array([1.0, 1.0, 0.0])
    p, f = case.params, case.fields
    o = load_oracle(case)
    o.create_cell_list()
    for op in ("hopkins_total.reset_pressure", "hopkins.compute_pressure", "hopkins_total.finalize_pressure",
               "hopkins_total.balance_of_momentum"):
        o.apply(op)
    d, r, mask = neighbours(f["x"], case.h)
    g = p["gamma"]
    h_ij = 0.5 * (f["h"][:, None] + f["h"][None, :])
    P = ((mask * f["m"][None, :] * (f["A"] ** (1 / g))[None, :] * wendland2(h_ij, r)).sum(1)) ** g
    assert rel_err(o.field("P"), P) < 1e-12
    v_pq = f["v"][:, None, :] - f["v"][None, :, :]
    dot = (d * v_pq).sum(-1)
    prefac = f["m"][None, :] * (f["A"][:, None] * f["A"][None, :]) ** (1 / g)
    e = 1.0 - 2.0 / g
    Pf = np.maximum(p["P_floor"], P)
    fc = -prefac * ((Pf ** e)[:, None] * rDwendland2(f["h"][:, None], r) + (Pf ** e)[None, :] * rDwendland2(f["h"][None, :], r))
    rfl = np.maximum(f["rho"], p["rho_floor"])
    cs = np.sqrt(g * P / rfl)
    mu = h_ij * dot / (r * r + p["eps"] * h_ij * h_ij)
    pi_ij = (-p["alpha"] * 0.5 * (cs[:, None] + cs[None, :]) * mu + p["beta"] * mu * mu) / (0.5 * (rfl[:, None] + rfl[None, :]))
    fv = np.where(dot < 0.0, -f["m"][None, :] * pi_ij * rDwendland2(h_ij, r), 0.0)
    Dv = ((mask * (fc + fv))[:, :, None] * d).sum(1)
    assert rel_err(o.field("Dv"), Dv) < 1e-12


# ---------------------------------------------------------------------------------------------
# Whole runs: 50 x verlet_step! with create_cell_list!'s removal, restated independently
# ---------------------------------------------------------------------------------------------
def numpy_remove_outside(case, f):
    """create_cell_list!'s removal, src/core.jl:60-81: particles outside the bounding box (closed
    intervals, geometry.jl:24-30) are overwritten, in DESCENDING index order, by the particles
    counted from the end, then the vector is shortened."""
    x = f["x"]
    lo, hi = np.asarray(case.box_min), np.asarray(case.box_max)
    inside = np.all((lo[None, :] <= x) & (x <= hi[None, :]), axis=1)
    removal = np.nonzero(~inside)[0][::-1]
    n = len(x)
    for i, idx in enumerate(removal, start=1):
        for a in f.values():
            a[idx] = a[n - i]
    keep = n - len(removal)
    for k in list(f):
        f[k] = f[k][:keep].copy()
    return len(removal)


def numpy_wcsph_run(case, nsteps):
    """nsteps of verlet_step! (wcsph_perturbed_witch.jl:309-332) with brute-force neighbours; the
    removal is create_cell_list!'s (:313), between the drift and the sums"""
    f = {k: v.copy() for k, v in case.fields.items()}
    removed = 0
    for _ in range(nsteps):
        numpy_step_kick_drift(f, case.params)
        removed += numpy_remove_outside(case, f)
        for k in ("P", "P_p", "Dv_last"):
            f.pop(k, None)
        numpy_step_sums(f, case.params, case.h, case.dim == 3)
    return f, removed


@pytest.mark.parametrize("make,nsteps", [
    (lambda: cases.mountain_wave_2d(n_y=12.0, dom_length=30e3, h_m=3000.0, a=6e3, U=25.0), 50),
    (lambda: cases.bell_hill_3d(10, 8, 6, h_m=3000.0, a=4e3, U=25.0), 50),
])
def test_fifty_steps_with_removal_against_numpy_restatement(make, nsteps):
    """A second, independent whole-run restatement (brute-force neighbours, no cell list, numpy
    formulas) against the C oracle over 50 steps, with particles leaving the box on the way — the
    oracle's cell list, swap-from-end removal (core.jl:72-81) and step sequence are what would
    break this by many orders of magnitude.  Summation order differs: agreement to rounding."""
    case = make()
    rng = np.random.default_rng(3)
    fl = np.nonzero(case.fields["type"] == 0.0)[0]
    # a handful of fluid particles shot through the top fence: they leave within the run
    top = fl[np.argsort(case.fields["x"][fl, 1])[-6:]]
    case.fields["v"][top, 1] = 10000.0 + 800.0 * rng.uniform(size=len(top))
    o = load_oracle(case)
    o.create_cell_list()
    o.step("wcsph", nsteps)
    ref, removed = numpy_wcsph_run(case, nsteps)
    assert removed > 0 and len(o) == case.n - removed == len(ref["m"])
    for name in ("x", "v", "rho", "h", "m", "type"):
        assert rel_err(o.field(name), ref[name]) < 1e-9, name


def numpy_aflow_step(f, p, h):
    """one verlet_step! of src/legacy/adiabatic_flow_witch.jl:231-243 (without add_new_particles!, which
    the device tests cover), written from the Julia source with k-d tree neighbours"""
    FLUID = p["fluid"]
    cv = p["cp"] - p["R_mass"]
    gam = p["gamma"]
    ey = np.array([0.0, 1.0, 0.0])
    sn = np.sin(np.pi / 2 * (1 - (p["z_t"] - p["z_b"]) / p["z_b"]))

    def accelerate():                                    # :225-229, damping_structure :210-216
        fl = f["type"] == FLUID
        damp = np.where(f["x"][:, 1] >= p["z_t"] - p["z_b"], p["gamma_r"] * sn ** 2, 0.0)
        dv = 0.5 * p["dt"] * (f["Dv"] - p["g"] * ey[None, :] - damp[:, None] * ey[None, :])
        f["v"] = np.where(fl[:, None], f["v"] + dv, f["v"])

    accelerate()
    fl = f["type"] == FLUID
    f["Dv"] = np.zeros_like(f["Dv"])                     # move! :217-223
    f["x"] = np.where(fl[:, None], f["x"] + p["dt"] * f["v"], f["x"])
    f["rho"] = np.where(fl, 0.0, f["rho"])
    n = len(f["m"])
    i, j, d, r = neighbour_pairs(f["x"], h)
    both = fl[i] & fl[j]
    # find_density! with self = true (:159-163, :238): the particle itself contributes wendland2(h, 0)
    rho_sum = np.bincount(i[both], weights=(f["m"][j] * wendland2(h, r))[both], minlength=n)
    f["rho"] = f["rho"] + rho_sum + np.where(fl, f["m"] * wendland2(h, 0.0), 0.0)
    f["s"] = np.where(fl, f["S"] * f["rho"] / f["m"], f["s"])                        # find_s! :165-169
    with np.errstate(all="ignore"):
        T = f["rho"] ** (gam - 1.0) * np.exp(f["s"] / (f["rho"] * cv)) / (cv * (gam - 1.0))   # find_pressure! :171-176
    f["T"] = np.where(fl, T, f["T"])
    f["P"] = np.where(fl, p["R_mass"] * f["rho"] * f["T"], f["P"])
    ker = rDwendland2(h, r)
    u = f["v"][i] - f["v"][j]
    dot = (u * d).sum(-1)
    # entropy_production! :184-191
    dS = -4.0 * f["m"][i] * f["m"][j] * ker * p["mu"] / (f["T"][i] * f["rho"][i] * f["rho"][j]) * dot ** 2 \
        / (r * r + 0.01 * h * h) * p["dt"]
    f["S"] = f["S"] + np.bincount(i[both], weights=dS[both], minlength=n)
    # internal_force! :146-153
    kq = f["m"][j] * ker
    a1 = -kq * (f["P"][i] / f["rho"][i] ** 2 + f["P"][j] / f["rho"][j] ** 2)
    a2 = 8.0 * kq * p["mu"] / (f["rho"][i] * f["rho"][j]) * dot / (r * r + 0.01 * h * h)
    Dv = np.zeros_like(f["Dv"])
    for a in range(3):
        Dv[:, a] = np.bincount(i, weights=(a1 + a2) * d[:, a], minlength=n)
    f["Dv"] = Dv
    accelerate()


def test_adiabatic_flow_steps_against_numpy_restatement():
    """the oracle's adiabatic flow closures (find_density! with self, find_s!, find_pressure!,
    entropy_production!, internal_force!, move!, accelerate!) against an independent numpy restatement
    over three steps"""
    case = cases.aflow_2d(n_y=16.0, dom_length=24e3, h_m=4e3, a=4e3, U_max=40.0)
    o = load_oracle(case)
    assert o.create_cell_list() == case.n
    f = {k: np.array(v, dtype=np.float64, copy=True) for k, v in case.fields.items()}
    for op in ("aflow.find_density", "aflow.find_pressure", "aflow.find_pot_temp", "aflow.find_s", "flow.internal_force"):
        o.apply(op)                                      # what closes make_system(), :121-126
    for name in ("rho", "T", "P", "s", "Dv"):
        f[name] = o.field(name).copy()                   # start both from that state
    for step in range(3):
        # no INFLOW particle converts in this case (they do not move, :219): the step is the closures
        o.step("aflow", 1)
        numpy_aflow_step(f, case.params, case.h)
        for name in ("x", "v", "rho", "s", "T", "P", "S", "Dv"):
            assert rel_err(f[name], o.field(name)) <= 1e-11, (step, name)
