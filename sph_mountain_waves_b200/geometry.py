"""Host-side CSG shapes for input generation — mirrors src/geometry.jl.

Cold path (setup only).  Vectorised over point arrays of shape (N, 3); the
predicates follow the reference's floating-point expressions so that the set of
generated lattice points is the same.
Shapes covered: Box/Rectangle (geometry.jl:15-43), Circle (:50-68), Ellipse
(:76-98), Ball (:245-258), BooleanUnion/Intersection/Difference (:106-169),
Specification (:176-187), BoundaryLayer (:196-232).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np


class Shape:
    def is_inside(self, x: np.ndarray) -> np.ndarray:  # x: (N,3) -> bool (N,)
        raise NotImplementedError

    def boundarybox(self) -> "Box":
        raise NotImplementedError

    # geometry.jl:235-237
    def __add__(self, other):
        return BooleanUnion(self, other)

    def __sub__(self, other):
        return BooleanDifference(self, other)

    def __mul__(self, other):
        return BooleanIntersection(self, other)


@dataclass
class Box(Shape):
    """geometry.jl:15-34 — closed intervals."""
    x1_min: float
    x2_min: float
    x3_min: float
    x1_max: float
    x2_max: float
    x3_max: float

    def is_inside(self, x):
        return ((self.x1_min <= x[:, 0]) & (x[:, 0] <= self.x1_max) &
                (self.x2_min <= x[:, 1]) & (x[:, 1] <= self.x2_max) &
                (self.x3_min <= x[:, 2]) & (x[:, 2] <= self.x3_max))

    def boundarybox(self):
        return self


def Rectangle(x1_min, x2_min, x1_max, x2_max) -> Box:
    """geometry.jl:41-43"""
    return Box(float(x1_min), float(x2_min), 0.0, float(x1_max), float(x2_max), 0.0)


@dataclass
class Circle(Shape):
    """geometry.jl:50-68"""
    x1: float
    x2: float
    r: float

    def is_inside(self, x):
        return (x[:, 0] - self.x1) ** 2 + (x[:, 1] - self.x2) ** 2 <= self.r ** 2

    def boundarybox(self):
        return Rectangle(self.x1 - self.r, self.x2 - self.r, self.x1 + self.r, self.x2 + self.r)


@dataclass
class Ellipse(Shape):
    """geometry.jl:76-98"""
    x1: float
    x2: float
    r1: float
    r2: float

    def is_inside(self, x):
        return ((x[:, 0] - self.x1) / self.r1) ** 2 + ((x[:, 1] - self.x2) / self.r2) ** 2 <= 1

    def boundarybox(self):
        return Rectangle(self.x1 - self.r1, self.x2 - self.r2, self.x1 + self.r1, self.x2 + self.r2)


@dataclass
class Ball(Shape):
    """geometry.jl:245-258"""
    x1: float
    x2: float
    x3: float
    r: float

    def is_inside(self, x):
        return ((x[:, 0] - self.x1) ** 2 + (x[:, 1] - self.x2) ** 2 +
                (x[:, 2] - self.x3) ** 2 <= self.r ** 2)

    def boundarybox(self):
        return Box(self.x1 - self.r, self.x2 - self.r, self.x3 - self.r,
                   self.x1 + self.r, self.x2 + self.r, self.x3 + self.r)


@dataclass
class BooleanUnion(Shape):
    """geometry.jl:106-125"""
    s1: Shape
    s2: Shape

    def is_inside(self, x):
        return self.s1.is_inside(x) | self.s2.is_inside(x)

    def boundarybox(self):
        a, b = self.s1.boundarybox(), self.s2.boundarybox()
        return Box(min(a.x1_min, b.x1_min), min(a.x2_min, b.x2_min), min(a.x3_min, b.x3_min),
                   max(a.x1_max, b.x1_max), max(a.x2_max, b.x2_max), max(a.x3_max, b.x3_max))


@dataclass
class BooleanIntersection(Shape):
    """geometry.jl:132-151"""
    s1: Shape
    s2: Shape

    def is_inside(self, x):
        return self.s1.is_inside(x) & self.s2.is_inside(x)

    def boundarybox(self):
        a, b = self.s1.boundarybox(), self.s2.boundarybox()
        return Box(max(a.x1_min, b.x1_min), max(a.x2_min, b.x2_min), max(a.x3_min, b.x3_min),
                   min(a.x1_max, b.x1_max), min(a.x2_max, b.x2_max), min(a.x3_max, b.x3_max))


@dataclass
class BooleanDifference(Shape):
    """geometry.jl:158-169"""
    s1: Shape
    s2: Shape

    def is_inside(self, x):
        return self.s1.is_inside(x) & ~self.s2.is_inside(x)

    def boundarybox(self):
        return self.s1.boundarybox()


@dataclass
class Specification(Shape):
    """geometry.jl:176-187 — `f` takes the (N,3) array and returns a bool mask."""
    s: Shape
    f: Callable[[np.ndarray], np.ndarray]

    def is_inside(self, x):
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.asarray(self.f(x), dtype=bool) & self.s.is_inside(x)

    def boundarybox(self):
        return self.s.boundarybox()


@dataclass
class SlabClip(Shape):
    """Not in the reference: `s` restricted to the half-open x-slab [xa, xb).  Used to
    generate only one rank's share of a lattice (multi-GPU x-slabs) in the same
    plane-major order as the whole shape."""
    s: Shape
    xa: float
    xb: float

    def is_inside(self, x):
        return self.s.is_inside(x) & (x[:, 0] >= self.xa) & (x[:, 0] < self.xb)

    def boundarybox(self):
        b = self.s.boundarybox()
        return Box(max(b.x1_min, self.xa), b.x2_min, b.x3_min, min(b.x1_max, self.xb), b.x2_max, b.x3_max)


class BoundaryLayer(Shape):
    """geometry.jl:196-232 — points outside `s` with a lattice offset |dx| <= width
    that lands inside `s`."""

    def __init__(self, s: Shape, grid, width: float):
        from .grids import covering, dimension
        self.s = s
        self.dim = dimension(grid)
        self.dxs = covering(grid, Ball(0.0, 0.0, 0.0, float(width)))
        self.width = float(width)
        self._grid = grid

    def is_inside(self, x):
        out = np.zeros(len(x), dtype=bool)
        cand = np.nonzero(~self.s.is_inside(x))[0]
        if len(cand) == 0 or len(self.dxs) == 0:
            return out
        # only points within `width` of s's bounding box can qualify: cheap prefilter
        bb = self.s.boundarybox()
        w = self.width * (1 + 1e-9)
        xc = x[cand]
        near = ((xc[:, 0] >= bb.x1_min - w) & (xc[:, 0] <= bb.x1_max + w) &
                (xc[:, 1] >= bb.x2_min - w) & (xc[:, 1] <= bb.x2_max + w) &
                (xc[:, 2] >= bb.x3_min - w) & (xc[:, 2] <= bb.x3_max + w))
        cand = cand[near]
        if isinstance(self.s, Box) and self._separable():
            out[cand] = self._box_fast(x[cand])
            return out
        hit = np.zeros(len(cand), dtype=bool)
        chunk = max(1, 4_000_000 // max(1, len(cand)))
        for a in range(0, len(self.dxs), chunk):
            dx = self.dxs[a:a + chunk]
            todo = np.nonzero(~hit)[0]
            if len(todo) == 0:
                break
            pts = (x[cand[todo], None, :] + dx[None, :, :]).reshape(-1, 3)
            ins = self.s.is_inside(pts).reshape(len(todo), len(dx)).any(axis=1)
            hit[todo] |= ins
        out[cand] = hit
        return out

    # --- exact fast path for axis-aligned boxes on square/cubic lattices ---------
    def _separable(self):
        from .grids import CubicGrid, Squaregrid
        return isinstance(self._grid, (Squaregrid, CubicGrid))

    def _box_fast(self, x):
        """is_inside(x+dx, Box) is separable per axis and the offsets are i*dr per
        axis, so `exists dx` reduces to: the per-axis smallest |i| that lands in
        the interval, taken together, is one of the ball's offsets (the ball test
        is monotone in each |i|).  Evaluates the very same float expressions."""
        dr = self._grid.dr
        b = self.s
        lo = (b.x1_min, b.x2_min, b.x3_min)
        hi = (b.x1_max, b.x2_max, b.x3_max)
        imax = int(np.ceil(self.width / dr)) + 1
        steps = np.arange(-imax, imax + 1)
        steps = steps[np.argsort(np.abs(steps), kind="stable")]  # 0, -1, 1, -2, 2, ...
        need = []
        ndim = 3 if self.dim == 3 else 2
        for a in range(ndim):
            best = np.full(len(x), 10 ** 6, dtype=np.int64)
            for i in steps:
                y = x[:, a] + (i * dr)
                ok = (lo[a] <= y) & (y <= hi[a]) & (best == 10 ** 6)
                best[ok] = abs(int(i))
            need.append(best)
        if ndim == 2:
            need.append(np.zeros(len(x), dtype=np.int64))
        # ball membership of (|i|,|j|,|k|) with the reference's expression
        # (x-0)^2 + (y-0)^2 + (z-0)^2 <= r^2  (geometry.jl:252-254)
        n0, n1, n2 = need
        valid = (n0 < 10 ** 6) & (n1 < 10 ** 6) & (n2 < 10 ** 6)
        d0 = np.where(valid, n0, 0) * dr
        d1 = np.where(valid, n1, 0) * dr
        d2 = np.where(valid, n2, 0) * dr
        inball = (d0 - 0.0) ** 2 + (d1 - 0.0) ** 2 + (d2 - 0.0) ** 2 <= self.width ** 2
        # the offsets list itself is bounded by the covering's index range
        rng = int(np.ceil(self.width / dr))
        inball &= (n0 <= rng) & (n1 <= rng) & (n2 <= rng)
        return valid & inball

    def boundarybox(self):
        r = self.s.boundarybox()
        w = self.width
        if self.dim == 2:
            return Rectangle(r.x1_min - w, r.x2_min - w, r.x1_max + w, r.x2_max + w)
        return Box(r.x1_min - w, r.x2_min - w, r.x3_min - w, r.x1_max + w, r.x2_max + w, r.x3_max + w)


def is_inside(x, s: Shape):
    """geometry.jl:4-6 — accepts one point or an (N,3) array."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return bool(s.is_inside(x[None, :])[0])
    return s.is_inside(x)


def boundarybox(s: Shape) -> Box:
    return s.boundarybox()
