"""One rank of the open-box NCCL test (tests/test_gpu_slabs_nccl.py runs it under torchrun).

Particles leave the GLOBAL bounding box while the set is split into x-slabs: through the downstream
face (pulled in behind the fluid, so the fast flow carries particles out on the last rank) and
through the top (a sprinkle of fluid particles on every rank is shot upwards).  create_cell_list!
then moves the particles at the end of sys.particles into the vacated slots (src/core.jl:72-81),
which renumbers survivors anywhere in the domain, and because neighbours are visited in index order
those numbers decide the bits of every later sum.  With sphmw_comm_open_box the library gathers the
dropped indices after every exchange and replays that loop on every rank: rank 0 compares the
gathered result — indices included — bit for bit with the whole-domain run on its own GPU."""
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
    case = cases.bell_hill_3d(64, 10, 8, h_m=3000.0, a=8e3, U=60.0)
    dr = case.info["dr"]
    x, v, typ = case.fields["x"], case.fields["v"], case.fields["type"]
    fluid = typ == case.params["fluid"]
    # downstream face 0.4 dr behind the last fluid plane: the walls beyond it go at the first cell
    # list, fluid follows within a few steps
    xmax = x[fluid, 0].max()
    case.box_max = (xmax + 0.4 * dr, case.box_max[1], case.box_max[2])
    # every 97th fluid particle of the upper third is shot upwards fast enough to cross the wall
    # layer and the top of the box within the test (vertical axis: x[2] of the reference = column 1)
    top = case.box_max[1]
    dt = case.params["dt"]
    pick = np.nonzero(fluid & (x[:, 1] > 0.66 * x[fluid, 1].max()))[0][::97]
    v[pick, 1] = (top - x[pick, 1] + 0.5 * dr) / (np.arange(len(pick)) % 7 + 3) / dt
    return case


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    case = make_case()
    names = ("x", "v", "rho", "h")
    failures = []
    run = SlabRun.from_global_case(case, rank, world, device=local, flags=flags)
    run.use_library_transport(open_box=True)
    run.create_cell_list()
    run.step(1)
    run.step(nsteps - 1)
    gidx, got = run.owned_fields(names)
    parts = [None] * world
    dist.all_gather_object(parts, (gidx, got, run.comm_info()))
    if rank == 0:
        from util import load_gpu
        whole = load_gpu(case, flags=flags)
        n_first = whole.create_cell_list()
        whole.step(nsteps)
        n_end = len(whole)
        allg = np.concatenate([p[0] for p in parts])
        order = np.argsort(allg, kind="stable")
        lost = [p[2].get("lost") for p in parts]
        print(f"open box: {case.n} particles, {n_first} after the first cell list, {n_end} after {nsteps} steps; "
              f"lost per rank {lost}", flush=True)
        if not (n_end < n_first < case.n):
            failures.append(f"the case does not lose particles on the way ({case.n} -> {n_first} -> {n_end})")
        if sum(1 for k in lost if k) < min(world, 2):
            failures.append(f"particles left on fewer than two ranks: {lost}")
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
        print("SLAB_OPEN_BOX_OK" if not failures else "SLAB_OPEN_BOX_FAIL " + "; ".join(failures), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
