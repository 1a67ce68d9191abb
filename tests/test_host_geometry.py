"""Host lattice/CSG generators against the reference's geometry test
(sph_jl/tests/test_geometry.jl:18-48, :58-140): area of 2D shapes = particle count * dA
(rtol 1 %), volume of 3D shapes (rtol 3 %), over every lattice."""
import math

import numpy as np
import pytest

from sph_mountain_waves_b200 import (Ball, BoundaryLayer, Box, Circle, Ellipse, Grid, Rectangle, Specification,
                                     covering)
from sph_mountain_waves_b200.geometry import SlabClip


@pytest.mark.parametrize("symm", ["square", "hexagonal"])
def test_areas_2d(symm):
    dr = 1 / 200  # test_geometry.jl:8-9
    grid = Grid(dr, symm)
    shapes = {
        "rect": (Rectangle(-1.0, -0.5, 1.0, 0.5), 2.0),
        "circle": (Circle(0.1, 0.2, 0.7), math.pi * 0.49),
        "ellipse": (Ellipse(0.0, 0.0, 0.8, 0.4), math.pi * 0.32),
        "union": (Rectangle(-1.0, -0.5, 0.0, 0.5) + Rectangle(-0.5, -0.5, 1.0, 0.5), 2.0),
        "difference": (Rectangle(-1.0, -0.5, 1.0, 0.5) - Circle(0.0, 0.0, 0.3), 2.0 - math.pi * 0.09),
        "intersection": (Rectangle(-1.0, -0.5, 1.0, 0.5) * Circle(0.0, 0.0, 0.4), math.pi * 0.16),
        "specification": (Specification(Rectangle(-1.0, -0.5, 1.0, 0.5), lambda x: x[:, 0] > 0.0), 1.0),
    }
    for name, (s, area) in shapes.items():
        n = len(covering(grid, s))
        assert n * dr * dr == pytest.approx(area, rel=1e-2), name


def test_boundary_layer_area_and_fast_path():
    dr = 0.02
    width = 6 * dr
    inner = Rectangle(-1.0, -0.5, 1.0, 0.5)
    for symm in ("square", "hexagonal"):
        grid = Grid(dr, symm)
        bl = BoundaryLayer(inner, grid, width)
        n = len(covering(grid, bl))
        # ring of width w with rounded corners
        area = 2 * width * (2.0 + 1.0) + math.pi * width * width
        assert n * dr * dr == pytest.approx(area, rel=0.1), symm  # the layer is a whole number of lattice rows
    # the separable fast path for boxes equals the brute-force definition (geometry.jl:207-217)
    grid = Grid(dr, "square")
    bl = BoundaryLayer(inner, grid, width)
    pts = covering(grid, Rectangle(-1.3, -0.8, 1.3, 0.8))
    fast = bl.is_inside(pts)
    brute = np.zeros(len(pts), dtype=bool)
    out = ~inner.is_inside(pts)
    for dx in bl.dxs:
        brute |= out & inner.is_inside(pts + dx)
    assert np.array_equal(fast, brute)


@pytest.mark.parametrize("symm", ["cubic", "facecentered", "bodycentered", "diamond"])
def test_volumes_3d(symm):
    dr = 0.01
    grid = Grid(dr, symm)
    shapes = {
        "box": (Box(-1.0, -0.5, -0.5, 1.0, 0.5, 0.5), 2.0),
        "ball": (Ball(0.0, 0.0, 0.0, 0.7), 4 / 3 * math.pi * 0.343),
    }
    for name, (s, vol) in shapes.items():
        n = len(covering(grid, s))
        assert n * dr ** 3 == pytest.approx(vol, rel=3e-2), (symm, name)


def test_boundary_layer_3d_fast_path_equals_brute_force():
    dr = 0.1
    grid = Grid(dr, "cubic")
    inner = Box(-0.5, -0.3, -0.4, 0.5, 0.3, 0.4)
    bl = BoundaryLayer(inner, grid, 3 * dr)
    pts = covering(grid, Box(-1.0, -0.8, -0.9, 1.0, 0.8, 0.9))
    fast = bl.is_inside(pts)
    brute = np.zeros(len(pts), dtype=bool)
    out = ~inner.is_inside(pts)
    for dx in bl.dxs:
        brute |= out & inner.is_inside(pts + dx)
    assert np.array_equal(fast, brute)


def test_hex_lattice_order_and_offsets():
    """grids.jl:77-93 — i outer / j inner, truncating j % 2"""
    g = Grid(1.0, "hexagonal")
    pts = covering(g, Rectangle(-2.5, -2.5, 2.5, 2.5))
    a, b = g.a, g.b
    expect = []
    for i in range(int(math.floor(-2.5 / a)) - 1, int(math.ceil(2.5 / a)) + 1):
        for j in range(int(math.floor(-2.5 / b)), int(math.ceil(2.5 / b)) + 1):
            x1 = (i + math.fmod(j, 2) / 2) * a
            x2 = j * b
            if -2.5 <= x1 <= 2.5 and -2.5 <= x2 <= 2.5:
                expect.append((x1, x2, 0.0))
    assert np.array_equal(pts, np.array(expect))


def test_slab_clip_partitions_the_lattice():
    g = Grid(0.1, "cubic")
    s = Box(-1.0, 0.0, -0.5, 1.0, 0.6, 0.5)
    whole = covering(g, s)
    parts = [covering(g, SlabClip(s, a, b)) for a, b in ((-1e30, -0.33), (-0.33, 0.41), (0.41, 1e30))]
    assert np.array_equal(np.concatenate(parts), whole)
