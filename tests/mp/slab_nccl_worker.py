"""One rank of the world-size-N NCCL test (tests/test_gpu_slabs_nccl.py runs it under torchrun).

Every rank slices the same small 3D case, steps its slab through the REAL transports —
`library`: ncclSend/ncclRecv inside libsphmw (csrc/slab_comm.cu), overlapped with the interior
force pass; `python`: torch.distributed.batch_isend_irecv (slabs.exchange) with the second-stream
overlap — and rank 0 compares the gathered result, bit for bit, with the whole-domain run on its
own GPU (wcsph_perturbed_witch.jl:309-332 has no distributed path: any number of ranks must give
the reference's single-process sums)."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    scheme = sys.argv[3] if len(sys.argv) > 3 else "wcsph"
    if scheme == "wcsph":
        case = cases.bell_hill_3d(64, 10, 8, h_m=3000.0, a=8e3, U=60.0)   # fast flow: particles migrate
        names = ("x", "v", "rho", "h")
    else:
        # the pressure-entropy drivers: three ghost columns, records carry A (plain schedule)
        case = cases.hopkins_2d(scheme, n_y=16.0, dom_length=160e3, h_m=3000.0, a=10e3, U=250.0)
        names = ("x", "v", "rho", "h", "P", "A")
    failures = []
    for transport in ("library", "python"):
        run = SlabRun.from_global_case(case, rank, world, device=local, flags=flags)
        if transport == "library":
            run.use_library_transport()
        run.create_cell_list()
        run.step(1)            # plain schedule
        run.step(nsteps - 1)   # overlapped schedule
        gidx, got = run.owned_fields(names)
        info = run.comm_info() if transport == "library" else {}
        parts = [None] * world
        dist.all_gather_object(parts, (gidx, got, info))
        if rank == 0:
            from util import load_gpu
            whole = load_gpu(case, flags=flags)
            whole.create_cell_list()
            whole.step(nsteps)
            allg = np.concatenate([p[0] for p in parts])
            order = np.argsort(allg, kind="stable")
            if len(allg) != case.n or not np.array_equal(allg[order], np.arange(case.n)):
                failures.append(f"{transport}: owned sets do not partition the particles ({len(allg)} of {case.n})")
            else:
                for f in names:
                    arr = np.concatenate([p[1][f] for p in parts])[order]
                    if not np.array_equal(arr, whole.field(f)):
                        bad = int(np.sum(np.any(np.atleast_2d(arr.T != whole.field(f).T), axis=0)))
                        failures.append(f"{transport}: field {f} differs from the whole-domain run in {bad} particles")
            if transport == "library":
                ex = [p[2].get("exchanges") for p in parts]
                if any(e != nsteps + 1 for e in ex):
                    failures.append(f"library: expected {nsteps + 1} exchanges per rank, saw {ex}")
                print("comm_info", parts[0][2], flush=True)
            whole.close()
        run.sys.close()
        dist.barrier()
    if rank == 0:
        print("SLAB_NCCL_OK" if not failures else "SLAB_NCCL_FAIL " + "; ".join(failures), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
