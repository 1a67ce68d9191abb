"""The real multi-GPU path: world-size-2 (and 4) torchrun runs that step slab contexts through the
NCCL transports — inside the library (csrc/slab_comm.cu) and the torch.distributed one
(slabs.exchange) — with the overlapped schedule, and compare bitwise with the whole-domain run.
Needs as many GPUs as ranks (skipped otherwise: NCCL ranks cannot share a device)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,flags", [(2, 0), (2, 1), (4, 0)])
def test_nccl_transports_match_the_whole_domain_run(world, flags):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "mp" / "slab_nccl_worker.py"), "12", str(flags)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0 and "SLAB_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world,scheme", [(2, "hopkins"), (4, "hopkins_full")])
def test_nccl_transports_step_the_hopkins_drivers(world, scheme):
    """three pair passes per step on slabs with three ghost columns, through both transports
    (sphmw_step(ctx, "hopkins", n) on a context with a communicator; step_phase 0/1 around
    torch.distributed), bitwise equal to the whole-domain run"""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 30500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "mp" / "slab_nccl_worker.py"), "12", "0", scheme]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0 and "SLAB_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_open_box_removal_renumbers_like_the_reference(world):
    """particles leave the global box on several ranks (downstream face, top): the library transport
    replays create_cell_list!'s swap-from-end removal (src/core.jl:72-81) across ranks, so indices
    and every field match the whole-domain run bit for bit (tests/mp/slab_open_box_worker.py)"""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 31500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "mp" / "slab_open_box_worker.py"), "14", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0 and "SLAB_OPEN_BOX_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_flow_scheme_on_slabs_with_inflow_and_outflow(world):
    """the constant-U flow driver on x-slabs: inflow re-seeding (new global indices in the reference's
    order) and outflow by removal (swap-from-end renumbering), both across ranks, bitwise equal to the
    whole-domain run (tests/mp/slab_flow_worker.py)"""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 32500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "mp" / "slab_flow_worker.py"), "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0 and "SLAB_FLOW_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
