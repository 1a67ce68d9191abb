"""Shared helpers of the parity tests: load one Case into the CPU oracle and into
the device library, and compare fields."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
from sph_mountain_waves_b200 import cases


def load_oracle(case: cases.Case) -> O.OracleSystem:
    o = O.OracleSystem(case.box_min, case.box_max, case.h, case.params)
    o.append(case.fields)
    return o


def load_gpu(case: cases.Case, **kw):
    return cases.to_system(case, **kw)


def _clean_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """|a - b| with matching NaNs and matching infinities counted as equal"""
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & (a == b)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    return np.where(both_nan | same_inf, 0.0, d)


def rel_err(a: np.ndarray, b: np.ndarray, floor: float = 1e-2, min_scale: float = 0.0) -> float:
    """The north star's "within 1e-10 relative": the larger of
      * per component k of a vector field:  max_i |a_ik - b_ik| / max_i |b_ik|   (a vertical
        velocity of 1e-3 m/s is judged against the vertical velocities, not against U = 20 m/s), and
      * element-wise:  max_i |a_ik - b_ik| / max(|b_ik|, floor * max_i |b_ik|)   (every value down
        to `floor` of its component's scale carries the full relative bar; below that the floor
        keeps round-off around zero from counting as an infinite relative error).
    `min_scale`: a lower bound for every component's scale, for fields that are a residual of
    cancelling terms (the velocity of an atmosphere at rest is round-off around 0: its natural
    unit is g*dt, not its own magnitude).  0 if both are identically zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    a2 = a.reshape(a.shape[0], -1)
    b2 = b.reshape(b.shape[0], -1)
    worst = 0.0
    for k in range(a2.shape[1]):
        d = _clean_diff(a2[:, k], b2[:, k])
        m = float(np.max(d))
        if m == 0.0:
            continue
        fin = np.abs(np.where(np.isfinite(b2[:, k]), b2[:, k], 0.0))
        scale = max(float(np.max(fin)), float(min_scale))
        if not scale > 0:
            return np.inf
        worst = max(worst, m / scale, float(np.max(d / np.maximum(fin, floor * scale))))
    return worst


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """IEEE equality; +0 == -0 and NaN == NaN count as equal"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def n_mismatch(a, b) -> int:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))


def field_err(case, f: str, a, b, nsteps: int = 1) -> float:
    """rel_err with the natural lower bound for the scale of residual fields: a velocity is
    measured in units of at least g*dt per step (what gravity alone adds in a step — the velocity
    of an atmosphere at rest is the round-off of pressure gradient against buoyancy), an
    acceleration in units of at least g."""
    g = abs(float(case.params.get("g", 0.0) or 0.0))
    dt = abs(float(case.params.get("dt", 0.0) or 0.0))
    floor = {"v": g * dt * max(1, nsteps), "Dv": g}.get(f, 0.0)
    return rel_err(a, b, min_scale=floor)
