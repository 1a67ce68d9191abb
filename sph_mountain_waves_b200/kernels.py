"""Smoothing kernels — mirrors the exports of src/kernels.jl
(wendland1/2/3, Dwendland*, rDwendland*, DDwendland3, spline23/24, Dspline*, rDspline*).

`f(h, r)` accepts scalars or arrays and is evaluated ON THE DEVICE through
sphmw_kernel_eval (the same __device__ functions the pair kernels inline); there
is no host implementation.
"""
from __future__ import annotations

import numpy as np

from . import _capi

_NAMES = ["wendland1", "Dwendland1", "rDwendland1", "wendland2", "Dwendland2", "rDwendland2",
          "wendland3", "Dwendland3", "rDwendland3", "DDwendland3", "spline23", "Dspline23",
          "rDspline23", "spline24", "Dspline24", "rDspline24"]


def _make(name: str):
    def f(h, r, device: int = 0):
        hb, rb = np.broadcast_arrays(np.asarray(h, dtype=np.float64), np.asarray(r, dtype=np.float64))
        shape = hb.shape
        hh = np.ascontiguousarray(hb).ravel()
        rr = np.ascontiguousarray(rb).ravel()
        out = np.empty_like(hh)
        _capi.check(_capi.lib().sphmw_kernel_eval(name.encode(), _capi.ptr(hh), _capi.ptr(rr),
                                                  _capi.ptr(out), hh.size, device))
        return float(out[0]) if shape == () else out.reshape(shape)
    f.__name__ = name
    f.__doc__ = f"{name}(h, r) — src/kernels.jl, evaluated on the device"
    return f


for _n in _NAMES:
    globals()[_n] = _make(_n)
__all__ = list(_NAMES)
