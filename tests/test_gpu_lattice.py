"""Device-side input generation against the host generators that restate
src/grids.jl and src/geometry.jl: the same particles in the same (reference) order."""
import numpy as np
import pytest

from sph_mountain_waves_b200.schemes import wcsph_perturbed_witch as w
from util import rel_err

pytestmark = pytest.mark.gpu

CONFIGS = {
    "hex_static": dict(n_y=24.0, dom_length=70e3),                                   # shipped driver: h_m = a = 0
    "hex_witch": dict(n_y=30.0, dom_length=90e3, h_m=3000.0, a=10e3, U=20.0, mountain_type=w.MOUNTAIN),
    "square_witch": dict(n_y=20.0, dom_length=60e3, h_m=5000.0, a=8e3, U=10.0, mountain_type=w.MOUNTAIN,
                         grid="square"),
    "cubic_bell": dict(n_y=14.0, dim=3, grid="cubic", h_m=4000.0, a=8e3, U=20.0, mountain_type=w.MOUNTAIN),
}


def constants(name):
    k = w.Constants(**CONFIGS[name])
    if k.dim == 3:
        k.dom_length, k.dom_width = 40 * k.dr, 18 * k.dr
    return k


@pytest.mark.parametrize("name", list(CONFIGS))
def test_device_generation_equals_host_generation(gpu, name):
    k = constants(name)
    host = w.make_system(k, lean=True)
    ref = {f: np.concatenate(v) for f, v in host._staged.items()}
    dev = w.make_system_on_device(k)
    assert list(dev.group_counts) == list(host.group_counts)
    assert len(dev) == len(ref["x"])
    for f in ("x", "v", "h", "type", "rho_p"):
        assert np.array_equal(dev.field(f), ref[f]), f
    for f in ("rho", "m"):                      # CUDA exp vs libm exp
        assert rel_err(dev.field(f), ref[f]) < 1e-15, f
    # and the generated system steps like the uploaded one
    host.create_cell_list()
    dev.create_cell_list()
    assert np.array_equal(host.cell_keys(), dev.cell_keys())
    host.step(2)
    dev.step(2)
    assert rel_err(dev.field("rho"), host.field("rho")) < 1e-13


def test_device_generation_on_slabs_partitions_the_lattice(gpu):
    from sph_mountain_waves_b200.slabs import plan_slab
    k = constants("cubic_bell")
    whole = w.make_system_on_device(k)
    x_all = whole.field("x")
    b = whole.domain
    got = []
    counts = []
    for r in range(3):
        plan = plan_slab((b.x1_min, b.x2_min, b.x3_min), (b.x1_max, b.x2_max, b.x3_max), k.h0, r, 3)
        part = w.make_system_on_device(k, capacity=len(x_all), slab=(plan.lo, plan.hi))
        n = part.n_device
        buf = np.empty((3, n))
        from sph_mountain_waves_b200 import _capi
        _capi.check(_capi.lib().sphmw_download_raw(part.ctx, b"x", _capi.ptr(buf), n, 3))
        x = np.ascontiguousarray(buf.T)
        assert np.all(plan.owns(x[:, 0]))
        got.append(x)
        counts.append(part.group_counts)
    assert np.array_equal(np.sum(counts, axis=0), whole.group_counts)
    both = np.concatenate(got)
    assert len(both) == len(x_all)
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(both), key(x_all))
