"""Mirror of src/current/wcsph_perturbed_witch.jl — static/perturbed atmosphere above a
Witch-of-Agnesi mountain, WCSPH with Monaghan artificial viscosity.

The driver's module-level `const`s (:25-75) become a `Constants` dataclass so the
configurations of BASELINE.json can be expressed (dr, mountain, inflow speed, 3D
extrusion) without editing the file; the defaults are the reference's.  The
closures (:195-303) are `Operator`s of the device menu; `verlet_step` (:309-332)
is the same call sequence, `verlet_step_fused` the library's fused equivalent.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from ..geometry import BoundaryLayer, Box, Rectangle, SlabClip, Specification
from ..grids import Grid
from ..system import (Operator, ParticleSystem, ParticleType, apply, create_cell_list,
                      generate_particles, new_pvd_file, save_frame, save_pvd_file)

folder_name = "wcsph_perturbed_witch"
export_vars = ("v", "ρ", "P", "θ", "T", "type")  # :18

FLUID, WALL, MOUNTAIN = 0.0, 1.0, 2.0  # :69-71


@dataclass
class Constants:
    """wcsph_perturbed_witch.jl:25-75 (defaults are the reference's values)."""
    dom_height: float = 26e3
    dom_length: float = 400e3
    n_y: float = 75.0              # dr = dom_height / n_y  (:27)
    bc_layers: float = 6.0         # bc_width = 6 dr        (:28)
    h_m: float = 0.0               # hₘ (:29)
    a: float = 0.0                 # a  (:30)
    eta: float = 1.8               # η  (:33)
    rho0: float = 1.393            # ρ0 (:38)
    eps: float = 0.01              # ε  (:44)
    alpha: float = 0.1             # α  (:45)
    g: float = 9.81                # (:50)
    R_mass: float = 287.05         # (:51)
    z_b: float = 12e3              # zᵦ (:53)
    R_gas: float = 8.314           # (:57)
    T_bg: float = 250.0            # (:61)
    t_end: float = 20.0            # (:65)
    rho_floor: float = 1e-6        # (:74)
    P_floor: float = 1e-10         # (:75)
    # --- extensions for BASELINE configs 3-5 (not in the shipped driver) ---
    dim: int = 2
    dom_width: float = 0.0         # 3D: extent along x[3] (the hill is at z = 0)
    U: float = 0.0                 # uniform initial wind along x (isothermal_flow_witch.jl:33)
    mountain_type: float = FLUID   # the shipped driver generates the mountain as FLUID (:164)
    grid: str = "hexagonal"        # :154 ; 3D uses "cubic"

    # derived exactly as the driver writes them
    @property
    def dr(self): return self.dom_height / self.n_y
    @property
    def bc_width(self): return self.bc_layers * self.dr
    @property
    def h0(self): return self.eta * self.dr                     # :34
    @property
    def m0(self): return self.rho0 * self.dr * self.dr           # :39
    @property
    def c(self): return math.sqrt(65e3 * (7 / 5) / self.rho0)    # :40
    @property
    def nu(self): return 0.1 * self.h0 * self.c                  # :43
    @property
    def beta(self): return 2 * self.alpha                        # :46
    @property
    def N(self): return math.sqrt(0.0196)                        # :49
    @property
    def gamma_r(self): return 10 * self.N                        # :52
    @property
    def z_t(self): return self.dom_height                        # :54
    @property
    def cp(self): return 7 * self.R_mass / 2                     # :58
    @property
    def cv(self): return self.cp - self.R_mass                   # :59
    @property
    def gamma(self): return self.cp / self.cv                    # :60
    @property
    def dt(self): return 0.01 * self.h0 / self.c                 # :64
    @property
    def dt_frame(self): return self.t_end / 100                  # :66

    def params(self) -> Dict[str, float]:
        """the constants the closures capture, under libsphmw's parameter names"""
        return dict(dt=self.dt, g=self.g, c=self.c, gamma=self.gamma, alpha=self.alpha,
                    beta=self.beta, eps=self.eps, eta=self.eta, rho0=self.rho0,
                    R_mass=self.R_mass, R_gas=self.R_gas, T_bg=self.T_bg,
                    rho_floor=self.rho_floor, P_floor=self.P_floor, z_t=self.z_t, z_b=self.z_b,
                    gamma_r=self.gamma_r, fluid=FLUID)


# Particle struct, :83-102
Particle = ParticleType(
    "Particle",
    ("h", "x", "m", "v", "Dv", "ρ_bg", "ρ′", "ρ", "P_bg", "P′", "P", "θ_bg", "θ′", "θ",
     "T_bg", "T′", "T", "type"),
    scheme="wcsph")


# background state, :177-189
def background_density(k: Constants, y):
    return k.rho0 * np.exp(-y * k.g / (k.R_mass * k.T_bg))


def background_pressure(k: Constants, y):
    return k.R_mass * k.T_bg * background_density(k, y)


def background_pot_temperature(k: Constants, y):
    return k.T_bg * ((k.T_bg * k.R_gas * k.rho0) / background_pressure(k, y)) ** (2 / 7)


# the fields the step actually carries from one step to the next; everything else is
# recomputed inside verlet_step! before it is read
CORE_FIELDS = ("h", "x", "m", "v", "ρ′", "ρ", "type")
LeanParticle = ParticleType("LeanParticle", CORE_FIELDS, scheme="wcsph")


def particle_ctor(k: Constants, v, ptype: float, lean: bool = False):
    """≙ Particle(x, v, type) inner constructor, :103-145 (vectorised over x).
    `lean` returns only CORE_FIELDS (same values) for very large particle sets."""
    def mass(rho):
        # :143  obj.m = obj.ρ * dr * dr  (left to right); one more factor dr in 3D
        m = rho * k.dr * k.dr
        return m if k.dim == 2 else m * k.dr

    def ctor(x: np.ndarray):
        n = len(x)
        y = x[:, 1]
        rho_bg = background_density(k, y)
        if lean:
            vv = np.zeros((n, 3))
            vv[:] = v
            return {"h": np.full(n, k.h0), "x": x, "v": vv, "rho_p": np.zeros(n),
                    "rho": 0.0 + rho_bg, "type": np.full(n, ptype), "m": mass(0.0 + rho_bg)}
        P_bg = background_pressure(k, y)
        th_bg = background_pot_temperature(k, y)
        vv = np.zeros((n, 3))
        vv[:] = v
        return {
            "h": np.full(n, k.h0), "x": x, "v": vv, "Dv": np.zeros((n, 3)),
            "rho_bg": rho_bg, "rho_p": np.zeros(n), "rho": 0.0 + rho_bg,
            "P_bg": P_bg, "P_p": np.zeros(n), "P": 0.0 + P_bg,
            "theta_bg": th_bg, "theta_p": np.zeros(n), "theta": 0.0 + th_bg,
            "T_bg": np.full(n, k.T_bg), "T_p": np.zeros(n), "T": np.full(n, 0.0 + k.T_bg),
            "type": np.full(n, ptype),
            "m": mass(0.0 + rho_bg),
        }
    return ctor


def witch_profile(k: Constants, x, z=None):
    """:158 — (hₘ a²)/(x² + a²); 0/0 = NaN at x = 0 when hₘ = a = 0.
    3D: bell-shaped hill hₘ / (1 + (x²+z²)/a²)^{3/2}."""
    with np.errstate(invalid="ignore", divide="ignore"):
        if z is None:
            return (k.h_m * k.a ** 2) / (x ** 2 + k.a ** 2)
        t = 1 + (x ** 2 + z ** 2) / k.a ** 2
        return k.h_m / (t * np.sqrt(t))   # t^(3/2) in IEEE operations only (same bits on the device)


def make_system_on_device(k: Optional[Constants] = None, capacity: Optional[int] = None,
                          **sys_kw) -> ParticleSystem:
    """make_system() with the lattice, the CSG classification and the Particle constructor
    evaluated on the GPU (sphmw_generate_mountain_wave): same particles in the same order as
    `make_system(k, lean=True)`; rho and m agree to an ulp (CUDA exp vs libm exp)."""
    from .._capi import LatticeSetup
    k = k or Constants()
    w = k.bc_width
    if k.dim == 2:
        dmin, dmax = (-k.dom_length / 2.0, 0.0, 0.0), (k.dom_length / 2.0, k.dom_height, 0.0)
        bmin, bmax = (dmin[0] - w, dmin[1] - w, 0.0), (dmax[0] + w, dmax[1] + w, 0.0)
    else:
        dmin = (-k.dom_length / 2.0, 0.0, -k.dom_width / 2.0)
        dmax = (k.dom_length / 2.0, k.dom_height, k.dom_width / 2.0)
        bmin, bmax = tuple(v - w for v in dmin), tuple(v + w for v in dmax)
    grid_id = {"square": 0, "hexagonal": 1, "cubic": 2}[k.grid if k.dim == 2 else "cubic"]
    if capacity is None:
        if grid_id == 1:
            sx, sy = (4 / 3) ** (1 / 4) * k.dr, (3 / 4) ** (1 / 4) * k.dr
        else:
            sx = sy = k.dr
        ni = math.ceil(bmax[0] / sx) - math.floor(bmin[0] / sx) + 3
        nj = math.ceil(bmax[1] / sy) - math.floor(bmin[1] / sy) + 1
        nk = (math.ceil(bmax[2] / k.dr) - math.floor(bmin[2] / k.dr) + 1) if k.dim == 3 else 1
        capacity = ni * nj * nk + 1024
    from ..geometry import Box
    sys = ParticleSystem(LeanParticle, Box(bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]), k.h0,
                         params=k.params(), capacity=capacity, **sys_kw)
    su = LatticeSetup()
    su.grid, su.mountain, su.dr = grid_id, (1 if k.dim == 2 else 2), k.dr
    su.dom_min[:], su.dom_max[:] = dmin, dmax
    su.bc_width, su.h_m, su.a, su.U = w, k.h_m, k.a, k.U
    su.type_fluid, su.type_wall, su.type_mountain, su.h0 = FLUID, WALL, k.mountain_type, k.h0
    sys.group_counts = list(sys.generate_mountain_wave(su))
    sys.constants = k
    return sys


def make_system(k: Optional[Constants] = None, lean: bool = False, clip=None,
                **sys_kw) -> ParticleSystem:
    """≙ make_system(), :152-170 (2D) and its 3D extrusion.  `clip=(xa, xb)` generates
    only the lattice sites with xa <= x < xb (one rank's x-slab)."""
    k = k or Constants()
    if k.dim == 2:
        grid = Grid(k.dr, k.grid, K=1.0)  # :154
        domain = Rectangle(-k.dom_length / 2.0, 0.0, k.dom_length / 2.0, k.dom_height)  # :155
        mountain = Specification(domain, lambda x: x[:, 1] <= witch_profile(k, x[:, 0]))  # :158-159
    else:
        grid = Grid(k.dr, "cubic", K=1.0)
        domain = Box(-k.dom_length / 2.0, 0.0, -k.dom_width / 2.0,
                     k.dom_length / 2.0, k.dom_height, k.dom_width / 2.0)
        mountain = Specification(domain, lambda x: x[:, 1] <= witch_profile(k, x[:, 0], x[:, 2]))
    fence = BoundaryLayer(domain, grid, k.bc_width)  # :156
    sys = ParticleSystem(LeanParticle if lean else Particle, domain + fence, k.h0,
                         params=k.params(), **sys_kw)  # :161
    wind = np.array([k.U, 0.0, 0.0])
    cl = (lambda s: s) if clip is None else (lambda s: SlabClip(s, clip[0], clip[1]))
    counts = [
        generate_particles(sys, grid, cl(domain - mountain), particle_ctor(k, wind, FLUID, lean)),  # :162
        generate_particles(sys, grid, cl(fence), particle_ctor(k, 0.0, WALL, lean)),  # :163
        generate_particles(sys, grid, cl(mountain),
                           particle_ctor(k, wind if k.mountain_type == FLUID else 0.0,
                                         k.mountain_type, lean)),  # :164
    ]
    sys.constants = k
    sys.group_counts = counts
    return sys


# closures, :195-303
compute_pressure = Operator("wcsph.compute_pressure")
find_temperature = Operator("wcsph.find_temperature")
find_pot_temp = Operator("wcsph.find_pot_temp")
reset_density = Operator("wcsph.reset_density")
compute_density = Operator("wcsph.compute_density")
finalize_density = Operator("wcsph.finalize_density")
update_smoothing = Operator("wcsph.update_smoothing")
balance_of_momentum = Operator("wcsph.balance_of_momentum")
move = Operator("wcsph.move")
accelerate = Operator("wcsph.accelerate")


def verlet_step(sys: ParticleSystem):
    """≙ verlet_step!(sys), :309-332 — operator by operator."""
    apply(sys, accelerate)
    apply(sys, move)
    create_cell_list(sys)
    apply(sys, reset_density)
    apply(sys, compute_density)
    apply(sys, finalize_density)
    apply(sys, update_smoothing)
    create_cell_list(sys)
    apply(sys, compute_pressure)
    apply(sys, find_temperature)
    apply(sys, find_pot_temp)
    apply(sys, balance_of_momentum)
    apply(sys, accelerate)


def verlet_step_fused(sys: ParticleSystem, nsteps: int = 1):
    """the same step through sphmw_step (fused kernels, state stays in HBM)"""
    sys.step(nsteps, "wcsph")


def avg_velocity(sys: ParticleSystem) -> float:
    """:338-345"""
    return sys.reduce("avg_speed")


def max_velocity(sys: ParticleSystem) -> float:
    """:347-350"""
    return sys.reduce("max_speed")


def main(k: Optional[Constants] = None, outpath: str = "results/" + folder_name, nsteps: Optional[int] = None,
         fused: bool = True, verbose: bool = True):
    """≙ main(), :356-407 (without the plots)."""
    k = k or Constants()
    sys = make_system(k)
    create_cell_list(sys)  # :166
    out = new_pvd_file(outpath)
    save_frame(out, sys, *export_vars)
    nsteps = int(round(k.t_end / k.dt)) if nsteps is None else nsteps
    every = int(round(k.dt_frame / k.dt))  # :375
    hist = []
    for step in range(1, nsteps + 1):
        if fused:
            verlet_step_fused(sys)
        else:
            verlet_step(sys)
        if step % every == 0:
            t = step * k.dt
            u_avg, u_max = avg_velocity(sys), max_velocity(sys)
            hist.append((t, u_avg, u_max))
            if verbose:
                print(f"t = {t}\nnum. of particles = {len(sys)}\nu_avg = {u_avg}\nu_max = {u_max}")
            save_frame(out, sys, *export_vars)
    save_pvd_file(out)
    return sys, hist
