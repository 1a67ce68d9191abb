"""sph_mountain_waves_b200 — B200-native WCSPH hot path (cell list, pair sums, Verlet step)
behind the SmoothedParticles.jl API of moschehaus/sph-mountain-waves.

Everything that computes runs in libsphmw.so (hand-written CUDA for sm_100a, C ABI in
include/sphmw.h).  Importing the package does not need a GPU; any call that touches
particles does, and fails loudly without one (no CPU fallback).
"""
from ._capi import LIB_PATH, SphmwError, UnsupportedOperator  # noqa: F401
from .geometry import (Ball, BooleanDifference, BooleanIntersection, BooleanUnion, BoundaryLayer,  # noqa: F401
                       Box, Circle, Ellipse, Rectangle, Shape, Specification, boundarybox, is_inside)
from .grids import (BodycenteredGrid, CubicGrid, DiamondGrid, FacecenteredGrid, Grid, Hexagrid,  # noqa: F401
                    Squaregrid, covering, dimension)
from .system import (DataStorage, Operator, ParticleField, ParticleSystem, ParticleType, apply,  # noqa: F401
                     apply_binary, apply_unary, create_cell_list, generate_particles, new_pvd_file,
                     import_particles, op_menu, read_vtp, save_frame, save_pvd_file, write_vtp)
from . import kernels  # noqa: F401

__version__ = "0.1.0"
