"""The N > 1 host logic of the x-slab decomposition on CPU: world_size 2 and 3 over the
`gloo` backend.  Partition planning, ownership by cell column, the migrant/ghost record
protocol and the transport are the product's (sph_mountain_waves_b200/slabs.py); the
per-rank arithmetic is done by the CPU oracle, so the distributed result can be compared
BIT FOR BIT with the single-rank oracle (neighbours ordered by global particle index)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sph_mountain_waves_b200 import cases
from sph_mountain_waves_b200 import slabs as S

CARRIED = ("x", "v", "m", "h", "rho", "rho_p", "type")


class OracleSlabBackend:
    """pack/unpack in numpy, compute with the oracle on owned + ghost particles"""

    def __init__(self, case, plan):
        from oracle import oracle as O
        self.O = O
        self.case, self.plan = case, plan
        own = plan.owns(case.fields["x"][:, 0])
        self.own = {f: case.fields[f][own].copy() for f in CARRIED}
        self.gidx = np.nonzero(own)[0].astype(np.int64)
        self.ghost = {f: self.own[f][:0].copy() for f in CARRIED}
        self.ghost_idx = np.zeros(0, dtype=np.int64)

    def _system(self, fields, gidx):
        order = np.argsort(gidx, kind="stable")  # local index order == global index order
        o = self.O.OracleSystem(self.case.box_min, self.case.box_max, self.case.h, self.case.params)
        o.append({f: a[order] for f, a in fields.items()})
        return o, order

    def empty(self, n):
        return torch.empty((n, S.RECORD), dtype=torch.float64)

    def pre(self):
        o, order = self._system(self.own, self.gidx)
        o.apply("wcsph.accelerate")
        o.apply("wcsph.move")
        inv = np.argsort(order)
        self.own["x"], self.own["v"] = o.field("x")[inv], o.field("v")[inv]

    def _records(self, sel, kind):
        n = int(sel.sum())
        r = np.zeros((n, S.RECORD))
        r[:, 0:3] = self.own["x"][sel]
        r[:, 3:6] = self.own["v"][sel]
        for k, f in zip(range(6, 11), ("m", "h", "rho", "rho_p", "type")):
            r[:, k] = self.own[f][sel]
        r[:, 11] = self.gidx[sel]
        r[:, 12] = kind
        return r

    def pack(self):
        p = self.plan
        col = S.column_of(self.own["x"][:, 0], p.h, p.phase0)
        go_l, go_r = col < p.lo, col >= p.hi
        edge_l = ~go_l & ~go_r & (col < p.lo + S.GHOST_COLS)
        edge_r = ~go_l & ~go_r & (col >= p.hi - S.GHOST_COLS)
        send_l = np.concatenate([self._records(go_l, S.KIND_MIGRANT), self._records(edge_l, S.KIND_GHOST)])
        send_r = np.concatenate([self._records(go_r, S.KIND_MIGRANT), self._records(edge_r, S.KIND_GHOST)])
        # migrants stay here as ghosts for this step; last step's ghosts are dropped
        gone = go_l | go_r
        self.ghost = {f: self.own[f][gone].copy() for f in CARRIED}
        self.ghost_idx = self.gidx[gone].copy()
        self.own = {f: a[~gone] for f, a in self.own.items()}
        self.gidx = self.gidx[~gone]
        sl = torch.from_numpy(send_l) if p.has_left else None
        sr = torch.from_numpy(send_r) if p.has_right else None
        return sl, sr, int(go_l.sum()), int(go_r.sum())

    def unpack(self, recv, n_migrants):
        if recv is None or len(recv) == 0:
            return
        r = recv.numpy()
        mig = r[:, 12] == S.KIND_MIGRANT
        assert int(mig.sum()) == n_migrants
        for sel, dst, name in ((mig, self.own, "gidx"), (~mig, self.ghost, "ghost_idx")):
            rec = r[sel]
            dst["x"] = np.concatenate([dst["x"], rec[:, 0:3]])
            dst["v"] = np.concatenate([dst["v"], rec[:, 3:6]])
            for k, f in zip(range(6, 11), ("m", "h", "rho", "rho_p", "type")):
                dst[f] = np.concatenate([dst[f], rec[:, k]])
            setattr(self, name, np.concatenate([getattr(self, name), rec[:, 11].astype(np.int64)]))

    def build(self):
        pass

    def post(self):
        fields = {f: np.concatenate([self.own[f], self.ghost[f]]) for f in CARRIED}
        gidx = np.concatenate([self.gidx, self.ghost_idx])
        o, order = self._system(fields, gidx)
        assert o.create_cell_list() == len(gidx)
        for op in ("wcsph.reset_density", "wcsph.compute_density", "wcsph.finalize_density",
                   "wcsph.update_smoothing", "wcsph.compute_pressure", "wcsph.balance_of_momentum",
                   "wcsph.accelerate"):
            o.apply(op)
        inv = np.argsort(order)
        n = len(self.gidx)
        for f in ("v", "rho", "rho_p", "h"):
            self.own[f] = o.field(f)[inv][:n]

    def counts(self):
        return len(self.gidx) + len(self.ghost_idx), len(self.gidx)


def _worker(rank, world, port, nsteps, U, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = cases.mountain_wave_2d(n_y=12.0, dom_length=90e3, h_m=3000.0, a=10e3, U=U)
        plan = S.plan_slab(case.box_min, case.box_max, case.h, rank, world)
        be = OracleSlabBackend(case, plan)
        run = S.SlabRun(None, plan, case.n, be)
        n0 = be.counts()[1]
        run.create_cell_list()
        run.step(nsteps)
        got = [None] * world
        dist.all_gather_object(got, (be.gidx, {f: be.own[f] for f in ("x", "v", "rho", "h")}, n0, be.counts()[1]))
        if rank == 0:
            gidx = np.concatenate([g[0] for g in got])
            order = np.argsort(gidx)
            np.savez(out_path, gidx=gidx[order], n0=np.array([g[2] for g in got]), n1=np.array([g[3] for g in got]),
                     **{f: np.concatenate([g[1][f] for g in got])[order] for f in ("x", "v", "rho", "h")})
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,U", [(2, 20.0), (3, 300.0)])
def test_gloo_slabs_equal_single_rank_oracle(tmp_path, world, U):
    from oracle import oracle as O
    O.set_threads(2)
    nsteps = 30
    out = str(tmp_path / "slabs.npz")
    mp.spawn(_worker, args=(world, _free_port(), nsteps, U, out), nprocs=world, join=True)
    got = np.load(out)
    case = cases.mountain_wave_2d(n_y=12.0, dom_length=90e3, h_m=3000.0, a=10e3, U=U)
    o = O.OracleSystem(case.box_min, case.box_max, case.h, case.params)
    o.append(case.fields)
    o.create_cell_list()
    o.step("wcsph", nsteps)
    assert np.array_equal(got["gidx"], np.arange(case.n))
    for f in ("x", "v", "rho", "h"):
        assert np.array_equal(got[f], o.field(f)), f
    assert got["n0"].sum() == got["n1"].sum() == case.n
    if U > 100:
        assert not np.array_equal(got["n0"], got["n1"]), "no particle migrated: the test is too tame"
    O.set_threads(O.max_threads())


def test_split_columns_and_global_indices():
    assert S.split_columns(100, 4) == [(0, 25), (25, 50), (50, 75), (75, 100)]
    with pytest.raises(ValueError):
        S.split_columns(10, 4)
    w = np.ones(40)
    w[:10] = 5.0
    parts = S.split_columns(40, 2, w)
    assert parts[0][1] < 20 and parts[0][0] == 0 and parts[1][1] == 40
    gc = np.array([[3, 2, 1], [4, 1, 0]])
    assert S.global_indices(gc, 0).tolist() == [0, 1, 2, 7, 8, 10]
    assert S.global_indices(gc, 1).tolist() == [3, 4, 5, 6, 9]
