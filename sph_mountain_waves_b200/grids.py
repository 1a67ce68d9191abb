"""Lattice generators for initial particle positions — mirrors src/grids.jl.

Cold path (setup only).  `covering(grid, shape)` returns the lattice points inside
`shape` as an (N,3) float64 array in the reference's loop order (first index
outermost), which fixes the initial particle indices and therefore the
cell-internal summation order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .geometry import Shape


@dataclass
class Squaregrid:
    """grids.jl:50-68"""
    dr: float


class Hexagrid:
    """grids.jl:70-93"""

    def __init__(self, dr: float):
        self.dr = float(dr)
        self.a = (4 / 3) ** (1 / 4) * self.dr
        self.b = (3 / 4) ** (1 / 4) * self.dr


@dataclass
class CubicGrid:
    """grids.jl:176-196"""
    dr: float


@dataclass
class BodycenteredGrid:
    """grids.jl:198-225"""
    dr: float


@dataclass
class FacecenteredGrid:
    """grids.jl:227-262"""
    dr: float


@dataclass
class DiamondGrid:
    """grids.jl:264-291"""
    dr: float


def Grid(dr: float, symm: str, K: float = 1.0):
    """grids.jl:28-40.  The fork made `K` a required keyword (SURVEY.md §4 caveat);
    the mountain drivers pass K=1.0 (wcsph_perturbed_witch.jl:154), which is the
    default here."""
    symm = symm.lstrip(":")
    table = {"square": Squaregrid, "hexagonal": Hexagrid, "cubic": CubicGrid,
             "facecentered": FacecenteredGrid, "bodycentered": BodycenteredGrid,
             "diamond": DiamondGrid}
    if symm not in table:
        raise ValueError("Unsupported grid type: " + symm)
    return table[symm](float(dr))


def dimension(grid) -> int:
    """grids.jl:42-48"""
    return 2 if isinstance(grid, (Squaregrid, Hexagrid)) else 3


def _range(lo: float, hi: float, step: float):
    return int(np.floor(lo / step)), int(np.ceil(hi / step))


def _filter(pts: np.ndarray, s: Shape, chunk: int = 8_000_000) -> np.ndarray:
    if len(pts) <= chunk:
        return pts[s.is_inside(pts)]
    keep = [pts[a:a + chunk][s.is_inside(pts[a:a + chunk])] for a in range(0, len(pts), chunk)]
    return np.concatenate(keep) if keep else pts[:0]


def _lattice3(ii, jj, kk, fx, fy, fz, s: Shape) -> np.ndarray:
    """points (fx(i), fy(j), fz(k)) for i outer, j, k inner; filtered slab by slab
    along i to bound memory."""
    out = []
    nj, nk = len(jj), len(kk)
    slab = max(1, 4_000_000 // max(1, nj * nk))
    yj = fy(jj)
    zk = fz(kk)
    for a in range(0, len(ii), slab):
        xi = fx(ii[a:a + slab])
        pts = np.empty((len(xi), nj, nk, 3))
        pts[..., 0] = xi[:, None, None]
        pts[..., 1] = yj[None, :, None]
        pts[..., 2] = zk[None, None, :]
        pts = pts.reshape(-1, 3)
        out.append(pts[s.is_inside(pts)])
    return np.concatenate(out) if out else np.zeros((0, 3))


def covering(grid, s: Shape) -> np.ndarray:
    box = s.boundarybox()
    if isinstance(grid, Squaregrid):
        i0, i1 = _range(box.x1_min, box.x1_max, grid.dr)
        j0, j1 = _range(box.x2_min, box.x2_max, grid.dr)
        ii = np.arange(i0, i1 + 1)
        jj = np.arange(j0, j1 + 1)
        return _lattice3(ii, jj, np.zeros(1, dtype=np.int64), lambda i: i * grid.dr,
                         lambda j: j * grid.dr, lambda k: k * 0.0, s)
    if isinstance(grid, Hexagrid):
        i_min = int(np.floor(box.x1_min / grid.a)) - 1
        j_min = int(np.floor(box.x2_min / grid.b))
        i_max = int(np.ceil(box.x1_max / grid.a))
        j_max = int(np.ceil(box.x2_max / grid.b))
        ii = np.arange(i_min, i_max + 1)
        jj = np.arange(j_min, j_max + 1)
        # (i + (j % 2)/2) * a with Julia's truncating remainder (grids.jl:85)
        shift = np.fmod(jj, 2) / 2
        out = []
        slab = max(1, 4_000_000 // max(1, len(jj)))
        x2 = jj * grid.b
        for a in range(0, len(ii), slab):
            i = ii[a:a + slab]
            pts = np.zeros((len(i), len(jj), 3))
            pts[..., 0] = (i[:, None] + shift[None, :]) * grid.a
            pts[..., 1] = x2[None, :]
            pts = pts.reshape(-1, 3)
            out.append(pts[s.is_inside(pts)])
        return np.concatenate(out) if out else np.zeros((0, 3))
    if isinstance(grid, CubicGrid):
        i0, i1 = _range(box.x1_min, box.x1_max, grid.dr)
        j0, j1 = _range(box.x2_min, box.x2_max, grid.dr)
        k0, k1 = _range(box.x3_min, box.x3_max, grid.dr)
        f = lambda i: i * grid.dr
        return _lattice3(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), np.arange(k0, k1 + 1), f, f, f, s)
    if isinstance(grid, BodycenteredGrid):
        a = 2 ** (1 / 3) * grid.dr
        i0, i1 = _range(box.x1_min, box.x1_max, a)
        j0, j1 = _range(box.x2_min, box.x2_max, a)
        k0, k1 = _range(box.x3_min, box.x3_max, a)
        ii, jj, kk = np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), np.arange(k0, k1 + 1)
        f0 = lambda i: i * a
        f1 = lambda i: (i + 0.5) * a
        return np.concatenate([_lattice3(ii, jj, kk, f0, f0, f0, s), _lattice3(ii, jj, kk, f1, f1, f1, s)])
    if isinstance(grid, FacecenteredGrid):
        a = 4 ** (1 / 3) * grid.dr
        i0, i1 = _range(box.x1_min, box.x1_max, a)
        j0, j1 = _range(box.x2_min, box.x2_max, a)
        k0, k1 = _range(box.x3_min, box.x3_max, a)
        ii, jj, kk = np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), np.arange(k0, k1 + 1)
        f0 = lambda i: i * a
        f1 = lambda i: (i + 0.5) * a
        first = _lattice3(ii, jj, kk, f0, f0, f0, s)
        # second loop pushes three candidates per (i,j,k), interleaved (grids.jl:247-260)
        I, J, Kk = np.meshgrid(ii, jj, kk, indexing="ij")
        I, J, Kk = I.ravel(), J.ravel(), Kk.ravel()
        trip = np.empty((len(I), 3, 3))
        trip[:, 0] = np.stack([(I + 0.5) * a, (J + 0.5) * a, Kk * a], axis=1)
        trip[:, 1] = np.stack([(I + 0.5) * a, J * a, (Kk + 0.5) * a], axis=1)
        trip[:, 2] = np.stack([I * a, (J + 0.5) * a, (Kk + 0.5) * a], axis=1)
        return np.concatenate([first, _filter(trip.reshape(-1, 3), s)])
    if isinstance(grid, DiamondGrid):
        a = 0.5 * grid.dr
        i0, i1 = _range(box.x1_min, box.x1_max, a)
        j0, j1 = _range(box.x2_min, box.x2_max, a)
        k0, k1 = _range(box.x3_min, box.x3_max, a)
        I, J, Kk = np.meshgrid(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), np.arange(k0, k1 + 1),
                               indexing="ij")
        I, J, Kk = I.ravel(), J.ravel(), Kk.ravel()
        odd = (I % 2 != 0)
        same = (odd == (J % 2 != 0)) & ((J % 2 != 0) == (Kk % 2 != 0))
        sm = np.fmod(I + J + Kk, 4)
        sm = np.fmod(sm + 4, 4)
        sel = same & ((sm == 0) | (sm == 1))
        pts = np.stack([I[sel] * a, J[sel] * a, Kk[sel] * a], axis=1)
        return _filter(pts, s)
    raise TypeError(f"unsupported grid {type(grid).__name__}")
