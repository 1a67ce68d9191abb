"""The oracle's smoothing kernels against every property the reference's own test pins
(sph_jl/tests/test_kernels.jl:19-61): zero outside the support, finite at r = 0, unit
integral (Simpson, rtol 1e-2), derivative consistency (rtol 1e-2), rD = D/r (rtol 1e-2)."""
import numpy as np
import pytest

from oracle import oracle as O

TOL = 0.01
N = 1000


def simpson_rule(f, a, b, n=N):
    # test_kernels.jl:9-17 (as written there: i runs 1..n-1)
    I = 0.0
    h = (b - a) / n
    for i in range(1, n):
        _a = a + i * h
        _b = a + (i + 1) * h
        I += h / 6.0 * (f(_a) + 4.0 * f(0.5 * (_a + _b)) + f(_b))
    return I


@pytest.mark.parametrize("dim,f,Df,rDf", [
    (1, "wendland1", "Dwendland1", "rDwendland1"),
    (2, "wendland2", "Dwendland2", "rDwendland2"),
    (3, "wendland3", "Dwendland3", "rDwendland3"),
    (2, "spline23", "Dspline23", "rDspline23"),
    (2, "spline24", "Dspline24", "rDspline24"),
])
def test_local_ker(dim, f, Df, rDf):
    # test_kernels.jl:19-43
    h = 0.42
    F = lambda r: O.kernel(f, h, r)
    DF = lambda r: O.kernel(Df, h, r)
    RDF = lambda r: O.kernel(rDf, h, r)
    assert F(4.0) == 0.0
    assert np.isfinite(F(0.0))
    if dim == 1:
        integral = simpson_rule(lambda r: 2.0 * F(r), 0.0, h)
    elif dim == 2:
        integral = simpson_rule(lambda r: 2.0 * np.pi * r * F(r), 0.0, h)
    else:
        integral = simpson_rule(lambda r: 4.0 * np.pi * r * r * F(r), 0.0, h)
    assert integral == pytest.approx(1.0, rel=TOL)
    assert DF(4.0) == 0.0
    assert np.isfinite(DF(0.0))
    integral = simpson_rule(DF, 0.2, 0.3)
    diff = F(0.3) - F(0.2)
    assert integral == pytest.approx(diff, rel=0.01)
    assert RDF(4.0) == 0.0
    assert np.isfinite(RDF(0.0))
    assert RDF(0.1) == pytest.approx(DF(0.1) / 0.1, rel=TOL)


def test_wendland_closed_forms():
    """kernels.jl:108-195 against an independent numpy evaluation of the formulas"""
    h = np.linspace(0.3, 2.0, 7)[:, None]
    r = np.linspace(0.0, 2.5, 41)[None, :]
    x = r / h
    inside = x <= 1.0
    w2 = np.where(inside, 7 / np.pi * (1 - x) ** 4 * (1 + 4 * x) / h ** 2, 0.0)
    rd2 = np.where(inside, -140 / np.pi * (1 - x) ** 3 / h ** 4, 0.0)
    w3 = np.where(inside, 21 / (2 * np.pi) * (1 - x) ** 4 * (1 + 4 * x) / h ** 3, 0.0)
    rd3 = np.where(inside, -210 / np.pi * (1 - x) ** 3 / h ** 5, 0.0)
    for name, ref in (("wendland2", w2), ("rDwendland2", rd2), ("wendland3", w3), ("rDwendland3", rd3)):
        got = O.kernel(name, h, r)
        assert np.allclose(got, ref, rtol=1e-13, atol=0.0), name
    # r == h is inside the support (x > 1 rejects) and evaluates to exactly zero there
    assert O.kernel("wendland2", 0.5, 0.5) == 0.0
