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


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max |b|  (0 if both are identically zero)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan, 0.0, np.abs(a - b))
    same_inf = np.isinf(a) & (a == b)
    d = np.where(same_inf, 0.0, d)
    scale = np.nanmax(np.abs(np.where(np.isfinite(b), b, 0.0)))
    m = float(np.max(d))
    if m == 0.0:
        return 0.0
    return m / scale if scale > 0 else np.inf


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """IEEE equality; +0 == -0 and NaN == NaN count as equal"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def n_mismatch(a, b) -> int:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))
