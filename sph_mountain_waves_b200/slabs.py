"""x-slab domain decomposition of one particle set over the GPUs of a box
(one process per GPU; SURVEY.md §8e — the reference has no distributed path).

Rank g owns the cell columns [X_g, X_{g+1}) of the global neighbour grid and keeps
two ghost columns on each side.  Every step one halo exchange (torch.distributed
send/recv over NCCL/NVLink) moves, between x-adjacent ranks only,
  * migrants — particles whose cell column left the slab (full carried state), and
  * ghosts   — copies of the particles in the two outermost owned columns.
Ghost density is recomputed locally (the second ghost column makes the first one's
sums complete), so there is no second exchange before the force pass, and because
neighbours are visited in (cell, global index) order the FP64 sums are bit-identical
for any number of ranks.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import cases
from .schemes import wcsph_perturbed_witch as wpw


def split_columns(ncols: int, world: int, weights: Optional[np.ndarray] = None):
    """Column ranges [lo, hi) per rank; equal widths for a uniform lattice, or
    balanced by per-column particle counts when `weights` is given."""
    if weights is None:
        edges = [round(ncols * r / world) for r in range(world + 1)]
    else:
        c = np.concatenate([[0], np.cumsum(weights)])
        target = c[-1] * np.arange(world + 1) / world
        edges = [int(np.searchsorted(c, t, side="left")) for t in target]
        edges[0], edges[-1] = 0, ncols
    for r in range(world):
        if edges[r + 1] - edges[r] < 4:
            raise ValueError("a slab must own at least 4 cell columns")
    return [(edges[r], edges[r + 1]) for r in range(world)]


class SlabRun:
    """One rank's share of a mountain-wave run."""

    def __init__(self, sys, rank: int, world: int, n_global: int, export=("v", "ρ", "P", "θ", "T", "type")):
        self.sys = sys
        self.rank, self.world = rank, world
        self.n_global = n_global
        self.export = export
        self._pinned = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def bell_hill_3d(cls, nx, ny, nz, rank=0, world=1, device=0, stream=None, **kw):
        if world == 1:
            case = cases.bell_hill_3d(nx, ny, nz, lean=True, **kw)
            sys = cases.to_system(case, device=device, stream=stream)
            sys._flush()
            run = cls(sys, rank, world, case.n)
            run.case_info = case.info
            return run
        raise NotImplementedError("multi-rank slabs: see SlabRun.distributed")

    # ------------------------------------------------------------------ stepping
    @property
    def n_owned(self) -> int:
        return self.sys.n_device

    @property
    def n_resident(self) -> int:
        return self.sys.n_device

    def create_cell_list(self):
        self.sys.create_cell_list()

    def step(self, nsteps: int):
        self.sys.step(nsteps)

    # ------------------------------------------------------------------ end to end
    def e2e_cycle(self, cycles: int, barrier: Callable[[], None]):
        """Through the public API with HOST buffers: upload the carried state from pinned
        host memory, run one output interval of steps (the reference saves a frame every
        Int(round(dt_frame/dt)) steps, wcsph_perturbed_witch.jl:375), download the exported
        fields (x + export_vars, :18) back to pinned host memory."""
        import time

        import torch
        sys = self.sys
        n = sys.n_device
        dim3 = 3
        carried = [f for f in wpw.CORE_FIELDS]
        every = int(self.case_info.get("frame_every", 8))
        if self._pinned is None:
            self._pinned = {}
            for f in set(carried) | set(self.export) | {"x"}:
                from .system import FIELD_NCOMP, canonical
                nc = FIELD_NCOMP[canonical(f)]
                self._pinned[f] = torch.empty((nc, n) if nc == 3 else (n,), dtype=torch.float64,
                                              pin_memory=True)
            for f in carried:
                sys.download_ptr(f, self._pinned[f].data_ptr(), n)
        h2d = sum(self._pinned[f].numel() * 8 for f in carried)
        d2h = sum(self._pinned[f].numel() * 8 for f in set(self.export) | {"x"})
        barrier()
        t0 = time.perf_counter()
        for _ in range(cycles):
            for f in carried:
                sys.upload_ptr(f, self._pinned[f].data_ptr(), n)
            sys.create_cell_list(want_count=False)
            sys.step(every)
            for f in set(self.export) | {"x"}:
                sys.download_ptr(f, self._pinned[f].data_ptr(), n)
        barrier()
        dt = time.perf_counter() - t0
        return {"seconds": dt, "steps": cycles * every, "h2d_bytes": h2d * cycles, "d2h_bytes": d2h * cycles,
                "what": f"{cycles} x (upload {len(carried)} carried fields from pinned host, create_cell_list, "
                        f"{every} steps = one frame interval, download x + {len(self.export)} export fields)"}
