"""The BASELINE.json configurations as backend-neutral particle sets.

A `Case` is plain host data (bounding box, h, driver constants, SoA field arrays in
the reference's particle index order).  `to_system` loads it into the device
library; tests load the same arrays into the CPU oracle.  All cases are
lattice-initialised and deterministic (no RNG), as SURVEY.md §8d prescribes.

  C1  collapse_dry            sph_jl/examples/collapse_dry.jl:30-106
  C2  static atmosphere 2D    src/current/wcsph_perturbed_witch.jl with dr = 26 km/120
  C3  Witch of Agnesi 2D      same driver, dr = 26 km/510, hₘ = 100 m, a = 10 km, U = 20 m/s
  C4  bell hill 3D            3D extrusion (cubic lattice, wendland3), 64 M particles
  C5  scaling sweep           C4 geometry at other sizes
  collision_2d                sph_jl/tests/test_collision_2d.jl:14-62
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from .geometry import BoundaryLayer, Circle, Rectangle, Specification
from .grids import Grid, covering
from .schemes import wcsph_perturbed_witch as wpw


@dataclass
class Case:
    name: str
    scheme: str
    dim: int
    box_min: tuple
    box_max: tuple
    h: float
    params: Dict[str, float]
    fields: Dict[str, np.ndarray]
    info: Dict[str, float] = field(default_factory=dict)

    @property
    def n(self) -> int:
        return len(self.fields["x"])


def _from_system(name: str, sys, scheme: str) -> Case:
    b = sys.domain
    f = {k: np.concatenate(v) for k, v in sys._staged.items()}
    ph, lim = None, None
    return Case(name, scheme, 2 if b.x3_min == 0.0 and b.x3_max == 0.0 else 3,
                (b.x1_min, b.x2_min, b.x3_min), (b.x1_max, b.x2_max, b.x3_max), sys.h,
                dict(sys.params), f)


def mountain_wave_2d(n_y: float = 75.0, h_m: float = 0.0, a: float = 0.0, U: float = 0.0,
                     dom_length: float = 400e3, name: Optional[str] = None, **kw) -> Case:
    """wcsph_perturbed_witch.jl make_system() at resolution dr = 26 km / n_y."""
    mt = wpw.FLUID if h_m == 0.0 else wpw.MOUNTAIN
    k = wpw.Constants(n_y=n_y, h_m=h_m, a=a, U=U, dom_length=dom_length, mountain_type=mt, **kw)
    sys = wpw.make_system(k)
    c = _from_system(name or f"mountain_wave_2d_ny{n_y:g}", sys, "wcsph")
    c.info = dict(dr=k.dr, dt=k.dt, frame_every=int(round(k.dt_frame / k.dt)))
    return c


def static_atmosphere_2d(n_y: float = 120.0, **kw) -> Case:
    """BASELINE config 2 (C2): hydrostatic well-balance test, ~250 k particles."""
    return mountain_wave_2d(n_y=n_y, name="C2_static_atmosphere_2d", **kw)


def witch_2d(n_y: float = 510.0, **kw) -> Case:
    """BASELINE config 3 (C3): ~4 M particles, hₘ = 100 m, a = 10 km, U = 20 m/s
    (mountain/inflow parameters from src/legacy/adiabatic_flow_witch.jl:31-33)."""
    return mountain_wave_2d(n_y=n_y, h_m=100.0, a=10e3, U=20.0, name="C3_witch_2d", **kw)


def bell_hill_3d(nx: int, ny: int, nz: int, h_m: float = 100.0, a: float = 10e3, U: float = 20.0,
                 lean: bool = False, clip: Optional[tuple] = None, name: Optional[str] = None) -> Case:
    """BASELINE configs 4/5: 3D extrusion on a cubic lattice with nx*ny*nz fluid
    cells, dr = 26 km / ny, vertical axis x[2].  `lean` keeps only the fields the step
    carries; `clip=(xa, xb)` generates only the sites with xa <= x < xb (one rank's
    x-slab; `info["group_counts"]` then gives the sizes of the three generation groups
    so that ranks can agree on global particle indices)."""
    k = wpw.Constants(n_y=float(ny), h_m=h_m, a=a, U=U, dim=3, grid="cubic",
                      mountain_type=wpw.MOUNTAIN if h_m else wpw.FLUID)
    k.dom_length = nx * k.dr
    k.dom_width = nz * k.dr
    sys = wpw.make_system(k, lean=lean, clip=clip)
    c = _from_system(name or f"bell_hill_3d_{nx}x{ny}x{nz}", sys, "wcsph")
    c.info = dict(dr=k.dr, dt=k.dt, frame_every=int(round(k.dt_frame / k.dt)),
                  group_counts=tuple(sys.group_counts))
    return c


def hopkins_2d(variant: str = "hopkins", n_y: float = 20.0, dom_length: float = 60e3, **kw) -> Case:
    """The pressure-entropy drivers on the same lattice: `variant` = "hopkins"
    (src/current/hopkins_perturbed_witch.jl), "hopkins_full" (full_hopkins_perturbed_witch.jl,
    adds A_bg = P_bg / rho_bg^gamma, :198-203) or "hopkins_total" (hopkins_total_witch.jl).
    Their constructors add A = P / rho^gamma and write m = rho * dr^2
    (hopkins_perturbed_witch.jl:146-147, hopkins_total_witch.jl:118-119)."""
    c = mountain_wave_2d(n_y=n_y, dom_length=dom_length, name=f"{variant}_2d_ny{n_y:g}", **kw)
    gamma = c.params["gamma"]
    dr = c.info["dr"]
    f = c.fields
    if variant in ("hopkins", "hopkins_full"):
        f["m"] = f["rho"] * (dr * dr)
    else:
        f["m"] = f["rho"] * dr * dr
    f["A"] = f["P"] / f["rho"] ** gamma
    if variant == "hopkins_full":
        f["A_bg"] = f["P_bg"] / f["rho_bg"] ** gamma
    if variant == "hopkins_total":
        # 11-field particle (hopkins_total_witch.jl:83-95)
        for k in ("rho_bg", "rho_p", "P_bg", "P_p", "theta_bg", "theta_p", "T_bg", "T_p"):
            f.pop(k)
    c.scheme = variant
    return c


def flow_2d(n_y: float = 100.0, dom_length: float = 100e3, h_m: float = 13e3, a: float = 10e3,
            U_max: float = 20.0) -> Case:
    """The constant-U flow driver with inflow re-seeding and outflow by domain removal —
    src/legacy/isothermal_flow_witch.jl:24-134 (the spec of SURVEY §8 f2).  That file cannot be
    loaded in the reference (it includes a missing packing module); the packing pre-step
    (:120) is therefore skipped here, everything else of make_system() is reproduced:
    square lattice, FLUID / MOUNTAIN / INFLOW / WALL groups in that order, OUTFLOW particles
    filtered out (:122), initialize_system!, set_density!, find_pressure!, find_pot_temp!."""
    dom_height = 26e3
    dr = dom_height / n_y
    h = 1.8 * dr
    bc_width = 6 * dr
    rho0, mu = 1.393, 15.98e-6
    c = math.sqrt(65e3 * (7 / 5) / rho0)
    N = math.sqrt(0.0196)
    g, R_mass, R_gas = 9.81, 287.05, 8.314
    cp = 7 * R_gas / 2
    T = 250.0
    dt = 0.01 * h / c
    FLUID, INFLOW, OUTFLOW, WALL, MOUNTAIN = 0.0, 1.0, 2.0, 3.0, 4.0
    grid = Grid(dr, "square")
    L2 = dom_length / 2.0
    domain = Rectangle(-L2, 0.0, L2, dom_height)
    fence = BoundaryLayer(domain, grid, bc_width)
    ground = Specification(fence, lambda x: x[:, 1] < 0)
    sky = Specification(fence, lambda x: x[:, 1] > dom_height)
    wind = Specification(fence, lambda x: (x[:, 0] <= -L2) & (x[:, 1] >= 0) & (x[:, 1] <= dom_height))
    sink = Specification(fence, lambda x: (x[:, 0] >= L2) & (x[:, 1] >= 0) & (x[:, 1] <= dom_height))
    with np.errstate(invalid="ignore", divide="ignore"):
        mountain = Specification(domain, lambda x: x[:, 1] <= (h_m * a ** 2) / (x[:, 0] ** 2 + a ** 2))
        groups = [(covering(grid, domain - mountain), FLUID), (covering(grid, mountain), MOUNTAIN),
                  (covering(grid, wind), INFLOW), (covering(grid, sink), OUTFLOW),
                  (covering(grid, ground + sky), WALL)]
    x = np.concatenate([g_[0] for g_ in groups])
    typ = np.concatenate([np.full(len(g_[0]), g_[1]) for g_ in groups])
    keep = typ != OUTFLOW                       # :122 filter!
    x, typ = x[keep], typ[keep]
    n = len(x)
    # Particle constructor :72-82, then :124-128
    rho = rho0 * np.exp(-x[:, 1] * g / (R_mass * T))
    m = rho * (dr * dr)
    u = np.zeros((n, 3))
    u[(typ == FLUID) | (typ == INFLOW), 0] = U_max      # initialize_system! :140-146
    rho = rho0 * np.exp(-x[:, 1] * g / (R_mass * T))    # set_density!
    P = rho * R_mass * T                                # find_pressure! (Drho = 0)
    theta = T * ((T * R_gas * rho0) / P) ** (R_gas / cp)
    fields = {"x": x, "v": u, "Dv": np.zeros((n, 3)), "rho": rho, "Drho": np.zeros(n), "m": m, "P": P,
              "theta": theta, "type": typ}
    params = dict(dt=dt, g=g, c=c, rho0=rho0, R_mass=R_mass, R_gas=R_gas, T_bg=T, mu=mu, kh=h,
                  z_t=dom_height, z_b=12e3, gamma_r=10 * N, fluid=FLUID, inflow=INFLOW, U_max=U_max, cp=cp,
                  bc_width=bc_width, x_inflow=-L2, dr=dr)
    box = (domain + fence).boundarybox()
    return Case("flow_2d", "flow", 2, (box.x1_min, box.x2_min, 0.0), (box.x1_max, box.x2_max, 0.0), h, params,
                fields, dict(dr=dr, dt=dt))


def aflow_2d(n_y: float = 100.0, dom_length: float = 100e3, h_m: float = 13e3, a: float = 10e3,
             U_max: float = 20.0, x_inflow: Optional[float] = None) -> Case:
    """The adiabatic variant of the flow driver — src/legacy/adiabatic_flow_witch.jl:24-128: the same
    geometry and particle groups as `flow_2d`, cp = 7 R_mass / 2 (:49), entropy S carried per particle.
    Fields are what the Particle constructor (:82-91) leaves plus initialize_system! (:134-140); the
    operator calls that close make_system() (:121-126: find_density!, find_pressure!, find_pot_temp!,
    find_s!, internal_force!) are the caller's, through the operator menu ("aflow.*").  The packing
    pre-step (:118) is skipped as in `flow_2d`.  `x_inflow` moves the line behind which INFLOW particles
    convert (the driver: -dom_length/2), so that a short test meets conversions."""
    dom_height = 26e3
    dr = dom_height / n_y
    h = 1.8 * dr
    bc_width = 6 * dr
    rho0, mu = 1.393, 15.98e-6
    c = math.sqrt(65e3 * (7 / 5) / rho0)
    N = math.sqrt(0.0196)
    g, R_mass, R_gas = 9.81, 287.05, 8.314
    cp = 7 * R_mass / 2
    cv = cp - R_mass
    gamma = cp / cv
    T0 = 250.0
    dt = 0.01 * h / c
    FLUID, INFLOW, OUTFLOW, WALL, MOUNTAIN = 0.0, 1.0, 2.0, 3.0, 4.0
    grid = Grid(dr, "square")
    L2 = dom_length / 2.0
    domain = Rectangle(-L2, 0.0, L2, dom_height)
    fence = BoundaryLayer(domain, grid, bc_width)
    ground = Specification(fence, lambda x: x[:, 1] < 0)
    sky = Specification(fence, lambda x: x[:, 1] > dom_height)
    wind = Specification(fence, lambda x: (x[:, 0] <= -L2) & (x[:, 1] >= 0) & (x[:, 1] <= dom_height))
    sink = Specification(fence, lambda x: (x[:, 0] >= L2) & (x[:, 1] >= 0) & (x[:, 1] <= dom_height))
    with np.errstate(invalid="ignore", divide="ignore"):
        mountain = Specification(domain, lambda x: x[:, 1] <= (h_m * a ** 2) / (x[:, 0] ** 2 + a ** 2))
        groups = [(covering(grid, domain - mountain), FLUID), (covering(grid, mountain), MOUNTAIN),
                  (covering(grid, wind), INFLOW), (covering(grid, sink), OUTFLOW),
                  (covering(grid, ground + sky), WALL)]
    x = np.concatenate([g_[0] for g_ in groups])
    typ = np.concatenate([np.full(len(g_[0]), g_[1]) for g_ in groups])
    keep = typ != OUTFLOW                       # :119 filter!
    x, typ = x[keep], typ[keep]
    n = len(x)
    # Particle constructor :82-91 (numpy's exp/log/power stand in for libm's: the tests compare the
    # device with the oracle on THESE inputs, and check the constructor itself through add_new_particles)
    T = np.full(n, T0)
    rho = rho0 * np.exp(-x[:, 1] * g / (R_mass * T))
    m = rho * dr ** 2
    P = R_mass * T * rho
    b = (T0 * R_gas * rho0) / P
    theta = T * (b * b) ** (1 / 7)
    S = m * cv * np.log((cv * T * (gamma - 1)) / (gamma * rho ** (gamma - 1)))
    u = np.zeros((n, 3))
    u[(typ == FLUID) | (typ == INFLOW), 0] = U_max      # initialize_system! :134-140
    fields = {"x": x, "v": u, "Dv": np.zeros((n, 3)), "rho": rho, "m": m, "P": P, "theta": theta, "T": T,
              "S": S, "s": np.zeros(n), "type": typ}
    params = dict(dt=dt, g=g, c=c, rho0=rho0, R_mass=R_mass, R_gas=R_gas, T_bg=T0, mu=mu, kh=h, gamma=gamma,
                  z_t=dom_height, z_b=12e3, gamma_r=10 * N, fluid=FLUID, inflow=INFLOW, U_max=U_max, cp=cp,
                  bc_width=bc_width, x_inflow=-L2 if x_inflow is None else x_inflow, dr=dr)
    box = (domain + fence).boundarybox()
    return Case("aflow_2d", "aflow", 2, (box.x1_min, box.x2_min, 0.0), (box.x1_max, box.x2_max, 0.0), h, params,
                fields, dict(dr=dr, dt=dt))


def collapse_dry(dr: float = 1.5e-2) -> Case:
    """BASELINE config 1 (C1) — sph_jl/examples/collapse_dry.jl:30-106."""
    # :42-62
    h = 3.0 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 2
    c = 50.0
    gy = -7.0 * 1.0  # g = -7.0*VECY
    mu = 8.4e-4
    nu = 1.0e-6
    water_column_width = 1.0
    water_column_height = 2.0
    box_height = 3.0
    box_width = 4.0
    wall_width = 2.5 * dr
    dt = 0.1 * h / c
    FLUID, WALL = 0.0, 1.0
    # :88-106
    grid = Grid(dr, "hexagonal")
    box = Rectangle(0.0, 0.0, box_width, box_height)
    fluid = Rectangle(0.0, 0.0, water_column_width, water_column_height)
    walls = BoundaryLayer(box, grid, wall_width)
    walls = Specification(walls, lambda x: x[:, 1] < box_height)
    domain = (box + walls).boundarybox()
    xf = covering(grid, fluid)
    xw = covering(grid, walls)
    x = np.concatenate([xf, xw])
    n = len(x)
    typ = np.concatenate([np.full(len(xf), FLUID), np.full(len(xw), WALL)])
    P = rho0 * gy * (x[:, 1] - water_column_height)
    rho = rho0 + P / c ** 2
    fields = {"x": x, "v": np.zeros((n, 3)), "Dv": np.zeros((n, 3)), "rho": rho,
              "Drho": np.zeros(n), "P": P, "type": typ}
    params = dict(dt=dt, c=c, rho0=rho0, m=m, nu=nu, mu=mu, gx=0.0, gy=gy, gz=0.0, kh=h, fluid=FLUID)
    return Case("C1_collapse_dry", "dambreak", 2, (domain.x1_min, domain.x2_min, 0.0),
                (domain.x1_max, domain.x2_max, 0.0), h, params, fields,
                dict(dr=dr, dt=dt, t_end=4.0))


def collision_2d(dr: float = 2.0e-2) -> Case:
    """sph_jl/tests/test_collision_2d.jl:14-62 — two colliding discs."""
    h = 2.4 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 2
    c = 20.0
    circ_rad = 0.4
    dom_len = dom_wid = 20.0
    deltaX, deltaY = 1.0, 0.2
    dt = 0.1 * h / c
    grid = Grid(dr, "square")
    circ1 = Circle(-0.5 * deltaX, -0.5 * deltaY, circ_rad)
    circ2 = Circle(0.5 * deltaX, 0.5 * deltaY, circ_rad)
    x1 = covering(grid, circ1)
    x2 = covering(grid, circ2)
    x = np.concatenate([x1, x2])
    n = len(x)
    v = np.zeros((n, 3))
    v[:len(x1), 0] = 1.0
    v[len(x1):, 0] = -1.0
    fields = {"x": x, "v": v, "Dv": np.zeros((n, 3)), "P": np.zeros(n), "rho": np.zeros(n),
              "rho0": np.zeros(n)}
    params = dict(dt=dt, c=c, rho0=rho0, m=m, kh=h)
    return Case("collision_2d", "collision", 2, (-0.5 * dom_len, -0.5 * dom_wid, 0.0),
                (0.5 * dom_len, 0.5 * dom_wid, 0.0), h, params, fields,
                dict(dr=dr, dt=dt, t_end=1.0))


# ---------------------------------------------------------------------------
_TYPES = {}


def particle_type_for(case: Case):
    from .system import ParticleType
    key = (case.scheme, tuple(sorted(case.fields)))
    if key not in _TYPES:
        _TYPES[key] = ParticleType("Particle_" + case.scheme, tuple(case.fields), scheme=case.scheme)
    return _TYPES[key]


def to_system(case: Case, **kw):
    """Load a case into a device ParticleSystem (libsphmw)."""
    from .geometry import Box
    from .system import ParticleSystem
    dom = Box(case.box_min[0], case.box_min[1], case.box_min[2], case.box_max[0], case.box_max[1],
              case.box_max[2])
    sys = ParticleSystem(particle_type_for(case), dom, case.h, params=case.params, **kw)
    sys.append(case.fields)
    return sys
