"""Mirror of src/utils/new_packing.jl — the damped pseudo-time relaxation ("packing") that
hopkins_total_witch.jl:142 runs before stepping.  Same call sequence (:64-140) over the device
operator menu (`packing.*`); the residuals (:82-89, :112-122) are evaluated every 10th
pseudo-step from downloaded fields, as the reference's serial loops do.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from ..system import Operator, ParticleSystem, apply, create_cell_list

reset_rho_pack = Operator("packing.reset_rho")
accumulate_rho_pack = Operator("packing.accumulate_rho")
balance_of_momentum_pack = Operator("packing.balance_of_momentum")
packing_accelerate = Operator("packing.accelerate")
packing_move = Operator("packing.move")


def packing_params(dt: float, c: float) -> dict:
    """new_packing.jl:1-3"""
    dt_pack = 1.0 * dt
    return dict(dt_pack=dt_pack, c_pack=2.0 * c, zeta_pack=1.0 * c / dt_pack)


def _residuals(sys: ParticleSystem):
    p = sys.params
    fluid = sys.field("type") == p["fluid"]
    y = sys.field("x")[fluid, 1]
    rho_t = p["rho0"] * np.exp(-y * p["g"] / (p["R_mass"] * p["T_bg"]))  # rho_target :18-20
    rho_err = math.sqrt(float(np.sum((sys.field("rho")[fluid] - rho_t) ** 2)))
    v = sys.field("v")[fluid]
    return rho_err, math.sqrt(float(np.sum(v * v)))


def packing(sys: ParticleSystem, abs_tol: float = 1e-3, rel_tol: float = 1e-2, maxSteps: int = 500,
            verbose: bool = False):
    """≙ packing!(sys; abs_tol, rel_tol, maxSteps) — new_packing.jl:64-140.  Returns the number
    of pseudo-steps taken."""
    for k, v in packing_params(sys.params["dt"], sys.params["c"]).items():
        sys.set_param(k, v)
    n = len(sys)
    sys.set_field("v", np.zeros((n, 3)))   # :70-73
    sys.set_field("Dv", np.zeros((n, 3)))
    create_cell_list(sys)
    apply(sys, reset_rho_pack)             # :76-77
    apply(sys, accumulate_rho_pack)
    rho_err0, _ = _residuals(sys)          # :80-87
    if verbose:
        print("---- PACKING INIT ----\nInitial density error =", rho_err0)
    k = 0
    while k < maxSteps:
        apply(sys, packing_accelerate)     # :96-98
        apply(sys, packing_move)
        create_cell_list(sys)
        apply(sys, reset_rho_pack)         # :101-102
        apply(sys, accumulate_rho_pack)
        apply(sys, balance_of_momentum_pack)  # :105-106
        apply(sys, packing_accelerate)
        if k % 10 == 0:                    # :109-127
            rho_err, v_norm = _residuals(sys)
            crit = abs_tol + rel_tol * rho_err0
            if verbose:
                print(f"packing step {k}: ρ_err = {rho_err}, |v| = {v_norm}, crit = {crit}")
            if rho_err < crit and v_norm < crit:
                break
        k += 1
    n = len(sys)
    sys.set_field("v", np.zeros((n, 3)))   # :133-136
    sys.set_field("Dv", np.zeros((n, 3)))
    return k
