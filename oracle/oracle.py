"""ctypes wrapper of oracle/libsph_oracle.so — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's hot path (oracle/sph_oracle.c, "parity
unpinned": see its header).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg may import this module; the product
(sph_mountain_waves_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Dict

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libsph_oracle.so"

FIELD_NCOMP = {
    "h": 1, "x": 3, "m": 1, "v": 3, "Dv": 3, "rho_bg": 1, "rho_p": 1, "rho": 1, "P_bg": 1,
    "P_p": 1, "P": 1, "theta_bg": 1, "theta_p": 1, "theta": 1, "T_bg": 1, "T_p": 1, "T": 1,
    "type": 1, "A": 1, "A_bg": 1, "Drho": 1, "rho0": 1, "S": 1, "s": 1,
}

_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "sph_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s", "-B" if force else "-s"], check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(str(LIB_PATH))
        P = C.c_void_p
        l.orc_create.restype = P
        l.orc_create.argtypes = [C.c_void_p, C.c_double]
        l.orc_destroy.argtypes = [P]
        l.orc_dim.argtypes = [P]
        l.orc_n.restype = C.c_int64
        l.orc_n.argtypes = [P]
        l.orc_key_max.restype = C.c_int64
        l.orc_key_max.argtypes = [P]
        l.orc_key_tables.argtypes = [P, C.c_void_p, C.c_void_p]
        l.orc_pair_count.restype = C.c_int64
        l.orc_pair_count.argtypes = [P]
        l.orc_set_threads.argtypes = [C.c_int]
        l.orc_set_param.argtypes = [P, C.c_char_p, C.c_double]
        l.orc_append.argtypes = [P, C.c_int64]
        l.orc_set_field.argtypes = [P, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64]
        l.orc_get_field.argtypes = [P, C.c_char_p, C.c_void_p]
        l.orc_create_cell_list.argtypes = [P]
        l.orc_cell_keys.argtypes = [P, C.c_void_p]
        l.orc_cell_entries.restype = C.c_int64
        l.orc_cell_entries.argtypes = [P, C.c_int64, C.c_void_p, C.c_int64]
        l.orc_pairs.restype = C.c_int64
        l.orc_pairs.argtypes = [P, C.c_void_p, C.c_void_p, C.c_int64]
        l.orc_apply.argtypes = [P, C.c_char_p, C.c_int]
        l.orc_step.argtypes = [P, C.c_char_p, C.c_int]
        l.orc_flow_add_new_particles.restype = C.c_int64
        l.orc_flow_add_new_particles.argtypes = [P]
        l.orc_aflow_add_new_particles.restype = C.c_int64
        l.orc_aflow_add_new_particles.argtypes = [P]
        l.orc_aflow_construct_all.restype = None
        l.orc_aflow_construct_all.argtypes = [P]
        for k in ("wendland1", "Dwendland1", "rDwendland1", "wendland2", "Dwendland2", "rDwendland2",
                  "wendland3", "Dwendland3", "rDwendland3", "DDwendland3", "spline23", "Dspline23",
                  "rDspline23", "spline24", "Dspline24", "rDspline24"):
            f = getattr(l, "orc_" + k)
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_double]
        _lib = l
    return _lib


def kernel(name: str, h, r):
    f = getattr(lib(), "orc_" + name)
    hb, rb = np.broadcast_arrays(np.asarray(h, dtype=np.float64), np.asarray(r, dtype=np.float64))
    out = np.array([f(float(a), float(b)) for a, b in zip(hb.ravel(), rb.ravel())])
    return float(out[0]) if hb.shape == () else out.reshape(hb.shape)


def set_threads(n: int):
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


class OracleSystem:
    """≙ ParticleSystem + create_cell_list! + apply! on the CPU (oracle)."""

    def __init__(self, box_min, box_max, h: float, params: Dict[str, float] | None = None):
        box = np.array(list(box_min) + list(box_max), dtype=np.float64)
        self._h = lib().orc_create(box.ctypes.data_as(C.c_void_p), float(h))
        if not self._h:
            raise AssertionError("invalid ParticleSystem declaration! (h must be a positive float)")
        for k, v in (params or {}).items():
            self.set_param(k, v)

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, name: str, v: float):
        if lib().orc_set_param(self._h, name.encode(), float(v)) != 0:
            raise KeyError(name)

    @property
    def n(self) -> int:
        return int(lib().orc_n(self._h))

    def __len__(self):
        return self.n

    @property
    def dim(self) -> int:
        return int(lib().orc_dim(self._h))

    def key_tables(self):
        ph = np.zeros(3, dtype=np.int64)
        lim = np.zeros(3, dtype=np.int64)
        lib().orc_key_tables(self._h, ph.ctypes.data_as(C.c_void_p), lim.ctypes.data_as(C.c_void_p))
        return tuple(int(v) for v in ph), tuple(int(v) for v in lim), int(lib().orc_key_max(self._h)), self.dim

    def append(self, fields: Dict[str, np.ndarray]):
        n_new = len(fields["x"])
        first = self.n
        lib().orc_append(self._h, n_new)
        for name, a in fields.items():
            if name not in FIELD_NCOMP:
                raise KeyError(name)
            self._set(name, np.asarray(a, dtype=np.float64), first, n_new)

    def _set(self, name, a, first, n):
        ncomp = FIELD_NCOMP[name]
        soa = np.ascontiguousarray(a.T if ncomp == 3 else a, dtype=np.float64)
        rc = lib().orc_set_field(self._h, name.encode(), soa.ctypes.data_as(C.c_void_p), first, n)
        if rc != 0:
            raise KeyError(name)

    def set_field(self, name: str, a):
        self._set(name, np.asarray(a, dtype=np.float64), 0, self.n)

    def field(self, name: str) -> np.ndarray:
        ncomp = FIELD_NCOMP[name]
        n = self.n
        buf = np.empty((ncomp, n) if ncomp == 3 else n, dtype=np.float64)
        if lib().orc_get_field(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p)) != 0:
            raise KeyError(name)
        return np.ascontiguousarray(buf.T) if ncomp == 3 else buf

    def create_cell_list(self) -> int:
        lib().orc_create_cell_list(self._h)
        return self.n

    def apply(self, op: str, self_: bool = False):
        if lib().orc_apply(self._h, op.encode(), 1 if self_ else 0) != 0:
            raise KeyError(f"oracle has no operator {op!r}")

    def step(self, scheme: str, nsteps: int = 1):
        if lib().orc_step(self._h, scheme.encode(), nsteps) != 0:
            raise KeyError(f"oracle has no scheme {scheme!r}")

    def flow_add_new_particles(self) -> int:
        return int(lib().orc_flow_add_new_particles(self._h))

    def aflow_add_new_particles(self) -> int:
        """add_new_particles! of src/legacy/adiabatic_flow_witch.jl:197-208"""
        return int(lib().orc_aflow_add_new_particles(self._h))

    def aflow_construct_all(self):
        """the adiabatic driver's Particle constructor (:82-91) on every particle: T, rho, m, P, theta, S from x"""
        lib().orc_aflow_construct_all(self._h)

    def cell_keys(self) -> np.ndarray:
        k = np.empty(self.n, dtype=np.int64)
        lib().orc_cell_keys(self._h, k.ctypes.data_as(C.c_void_p))
        return k

    def cell_entries(self, key: int) -> np.ndarray:
        n = lib().orc_cell_entries(self._h, key, None, 0)
        out = np.empty(n, dtype=np.int64)
        if n:
            lib().orc_cell_entries(self._h, key, out.ctypes.data_as(C.c_void_p), n)
        return out

    def pairs(self):
        n = lib().orc_pairs(self._h, None, None, 0)
        pi = np.empty(n, dtype=np.int64)
        pj = np.empty(n, dtype=np.int64)
        if n:
            lib().orc_pairs(self._h, pi.ctypes.data_as(C.c_void_p), pj.ctypes.data_as(C.c_void_p), n)
        return pi, pj

    def pair_count(self) -> int:
        return int(lib().orc_pair_count(self._h))
