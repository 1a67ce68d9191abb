"""One rank of the flow-on-slabs NCCL test (tests/test_gpu_slabs_nccl.py runs it under torchrun).

The constant-U flow driver (src/legacy/isothermal_flow_witch.jl) gains particles at the inflow
(add_new_particles!, :175-186: successor k of the step gets index N + k, k counting the converting
particles in index order) and loses them through the downstream face of the bounding box
(create_cell_list!'s swap-from-end removal, src/core.jl:72-81).  Both renumber particles across the
whole domain.  On x-slabs the library all-gathers the converting and the dropped indices and replays
both rules on every rank (csrc/slab_comm.cu, sphmw_comm_open_box): rank 0 compares the gathered result,
indices included, bit for bit with the whole-domain run on its own GPU."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from sph_mountain_waves_b200 import cases  # noqa: E402
from sph_mountain_waves_b200.slabs import SlabRun  # noqa: E402


def make_case():
    case = cases.flow_2d(n_y=20.0, dom_length=30e3, h_m=4e3, a=4e3, U_max=400.0)
    # downstream face 0.3 dr behind the fluid: particles leave within the test
    case.box_max = (15e3 + 0.3 * case.info["dr"], case.box_max[1], 0.0)
    return case


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    case = make_case()
    names = ("x", "v", "rho", "P", "m", "type")
    failures = []
    run = SlabRun.from_global_case(case, rank, world, device=local)
    run.use_library_transport(open_box=True)
    run.create_cell_list()
    n_start = sum(_all(run.n_owned))
    run.step(nsteps)
    gidx, got = run.owned_fields(names)
    parts = [None] * world
    dist.all_gather_object(parts, (gidx, got, run.comm_info()))
    if rank == 0:
        from util import load_gpu
        whole = load_gpu(case, capacity=2 * case.n)
        n_first = whole.create_cell_list()
        whole.step(nsteps, "flow")
        n_end = len(whole)
        allg = np.concatenate([p[0] for p in parts])
        order = np.argsort(allg, kind="stable")
        print(f"flow on slabs: {case.n} particles, {n_first} after the first cell list (slabs: {n_start}), "
              f"{n_end} after {nsteps} steps; lost per rank {[p[2].get('lost') for p in parts]}", flush=True)
        if n_start != n_first:
            failures.append(f"first cell list: {n_start} particles on the slabs, {n_first} on one GPU")
        if len(allg) != n_end or not np.array_equal(allg[order], np.arange(n_end)):
            failures.append(f"owned indices are not 0..{n_end - 1} ({len(allg)} particles)")
        else:
            for f in names:
                arr = np.concatenate([p[1][f] for p in parts])[order]
                if not np.array_equal(arr, whole.field(f)):
                    bad = int(np.sum(np.any(np.atleast_2d(arr.T != whole.field(f).T), axis=0)))
                    failures.append(f"field {f} differs from the whole-domain run in {bad} particles")
        whole.close()
    run.sys.close()
    dist.barrier()
    if rank == 0:
        print("SLAB_FLOW_OK" if not failures else "SLAB_FLOW_FAIL " + "; ".join(failures), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


def _all(v):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, int(v))
    return out


if __name__ == "__main__":
    main()
