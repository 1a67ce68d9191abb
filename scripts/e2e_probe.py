"""Where does an end-to-end cycle spend its time?  (diagnostic, not part of the bench)"""
import ctypes as C, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from sph_mountain_waves_b200 import _capi
from sph_mountain_waves_b200._capi import check
from sph_mountain_waves_b200.slabs import SlabRun
from sph_mountain_waves_b200.system import FIELD_NCOMP, canonical
from sph_mountain_waves_b200.schemes import wcsph_perturbed_witch as wpw

nx, ny, nz = (960, 75, 96) if len(sys.argv) < 2 else (1920, 150, 192)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    run = SlabRun.bell_hill_3d(nx, ny, nz, device=0, stream=stream.cuda_stream, flags=0, device_gen=True)
    run.create_cell_list()
    run.step(3)
    run.e2e_cycle(cycles=2, barrier=torch.cuda.synchronize)   # warm (allocations)
    s, lib = run.sys, _capi.lib()
    carried = list(wpw.CORE_FIELDS)
    out_fields = sorted(set(run.export) | {"x"})
    names = (C.c_char_p * len(out_fields))(*[canonical(f).encode() for f in out_fields])
    n0 = run._n0
    def T(label, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        print(f"{label:28s} {1e3*(time.perf_counter()-t0):8.2f} ms", flush=True)
    def prefetch():
        for f in carried:
            check(lib.sphmw_upload_async(s.ctx, canonical(f).encode(), C.c_void_p(run._pinned[f].data_ptr()), n0, FIELD_NCOMP[canonical(f)]))
    for rep in range(2):
        T("upload_async x7 (H2D)", prefetch)
        T("upload_commit", lambda: check(lib.sphmw_upload_commit(s.ctx)))
        T("create_cell_list", run.create_cell_list)
        T("step(1) first", lambda: run.step(1))
        T("step(15)", lambda: run.step(15))
        slot = C.c_int32()
        T("frame_capture (+D2H)", lambda: check(lib.sphmw_frame_capture(s.ctx, names, len(out_fields), C.byref(slot))))
        ptrs = (C.c_void_p * len(out_fields))(); n = C.c_int64()
        T("frame_wait", lambda: check(lib.sphmw_frame_wait(s.ctx, slot.value, ptrs, len(out_fields), C.byref(n))))
    # overlap probes
    T("step(8) alone", lambda: run.step(8))
    def up_and_step():
        prefetch(); run.step(8)
    T("prefetch + step(8)", up_and_step)
    check(lib.sphmw_upload_commit(s.ctx)); run.create_cell_list(); run.step(1)
    def cap_and_step():
        slot2 = C.c_int32()
        h0 = time.perf_counter()
        check(lib.sphmw_frame_capture(s.ctx, names, len(out_fields), C.byref(slot2)))
        h1 = time.perf_counter()
        run.step(8)
        h2 = time.perf_counter()
        print(f"   host time: capture call {1e3*(h1-h0):.2f} ms, step(8) call {1e3*(h2-h1):.2f} ms")
    T("capture + step(8)", cap_and_step)
    def both_and_step():
        slot2 = C.c_int32()
        check(lib.sphmw_frame_capture(s.ctx, names, len(out_fields), C.byref(slot2)))
        prefetch(); run.step(8)
    T("capture + prefetch + step(8)", both_and_step)
    check(lib.sphmw_upload_commit(s.ctx)); run.create_cell_list(); run.step(1)
    s.timing(True); s.timing_reset()
    t0 = time.perf_counter()
    r = run.e2e_cycle(cycles=4, barrier=torch.cuda.synchronize)
    print("e2e 4 cycles", r["seconds"], "s; per cycle", r["seconds"] / 4 * 1e3, "ms")
    rep = s.timing_report()
    tot = sum(v[0] for v in rep.values())
    print("kernel sum per cycle", tot / 4, "ms")
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"  {k:28s} {v[0]/4:9.2f} ms/cycle  calls/cycle {v[1]/4:.1f}")
