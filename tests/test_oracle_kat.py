"""Known-answer tests of the reference, restated against the CPU oracle.

Each test cites the reference test it restates (paths relative to /root/reference).  The
reference tests use `isapprox` (rtol = sqrt(eps)) or atol = 1e-10; we assert tighter.
Random inputs of the reference come from Julia's RNG and cannot be reproduced; the
*property* they test holds for any input, so NumPy's RNG is used instead.
"""
import itertools

import numpy as np
import pytest

from oracle import oracle as O

T1 = np.array([-3.14, 1.0, 3.0, 7.6, 12.8])        # test/test_interpolations.jl:9
T2 = np.array([-2.71, 1.41, 12.76, 50.2, 120.0])   # test/test_interpolations.jl:10


def mid(t):
    t = np.asarray(t)
    return t[:-1] + np.diff(t) / 2


# ---- get_idx (src/interpolation_utils.jl:104-134), edge cases of SURVEY.md §8a A4 -------
def test_get_idx_edge_cases():
    t = [0.0, 1.0, 2.0, 3.0]
    x = [-1.0, 0.0, 0.5, 1.0, 1.5, 3.0, 4.0]
    assert O.get_idx("linear", t, x).tolist() == [1, 1, 1, 1, 2, 3, 3]
    assert O.get_idx("constant", t, x).tolist() == [1, 1, 1, 2, 2, 3, 3]
    assert O.get_idx("bspline", t, x, degree=2).tolist() == [3, 3, 3, 4, 4, 5, 5]


def test_get_idx_isless_total_order():
    # Julia's searchsorted* use isless: -0.0 < +0.0, NaN sorts last.
    t = [-1.0, 0.0, 1.0]
    assert O.get_idx("constant", t, [-0.0, 0.0]).tolist() == [1, 2]
    assert O.get_idx("linear", t, [-0.0, 0.0, np.nan]).tolist() == [1, 1, 2]
    assert O.get_idx("constant", t, [np.nan, np.inf, -np.inf]).tolist() == [2, 2, 1]
    tm = [-1.0, -0.0, 1.0]
    assert O.get_idx("constant", tm, [-0.0, 0.0]).tolist() == [2, 2]
    assert O.get_idx("linear", tm, [-0.0, 0.0]).tolist() == [1, 2]


def test_expand_knots_default_and_custom():
    # src/interpolation_dimensions.jl:168-173, src/spline_utils.jl:1-20
    assert O.expand_knots([0.0, 1.0, 2.0], 2).tolist() == [0, 0, 0, 1, 2, 2, 2]
    assert O.expand_knots([0.0, 1.0, 2.0], 1, [2, 1, 2]).tolist() == [0, 0, 1, 2, 2]


# ---- test_globally_constant (test/test_interpolations.jl:5-50) ---------------------------
@pytest.mark.parametrize("case", ["linear", "bspline", "nurbs"])
def test_globally_constant(case):
    rng = np.random.default_rng(10)
    if case == "linear":
        kinds, degs, u, w, derivs = ["linear"] * 2, None, np.full((5, 5), 2.0), None, True       # :68
    elif case == "bspline":
        kinds, degs, u, w, derivs = ["bspline"] * 2, [2, 3], np.full((6, 7), 2.0), None, True     # :86-90
    else:
        kinds, degs, u, w, derivs = ["bspline"] * 2, [3, 1], np.full((7, 5), 2.0), rng.random((7, 5)), False  # :108-114
    for te in ((T1, T2), (mid(T1), mid(T2))):
        val = O.evaluate(kinds, [T1, T2], u, te, degrees=degs, weights=w, grid=True)
        assert val.shape == (len(te[0]), len(te[1]))
        np.testing.assert_allclose(val, 2.0, atol=1e-13, rtol=0)
        if derivs:
            for orders in ((1, 0), (0, 1)):
                d = O.evaluate(kinds, [T1, T2], u, te, degrees=degs, derivative_orders=orders, grid=True)
                np.testing.assert_allclose(d, 0.0, atol=1e-13, rtol=0)


# ---- test_analytic, Linear plane (test/test_interpolations.jl:52-82) ----------------------
def test_linear_reproduces_plane():
    rng = np.random.default_rng(1)
    t1, t2 = np.cumsum(rng.random(10)), np.cumsum(rng.random(10))
    f = lambda a, b: 3.0 + 2.3 * a - 4.7 * b
    u = f(t1[:, None], t2[None, :])
    for ts in ((t1, t2), (mid(t1), mid(t2))):
        for a, b in itertools.product(*ts):
            got = O.point(["linear"] * 2, [t1, t2], u, (a, b))
            assert got == pytest.approx(f(a, b), rel=1e-13, abs=1e-13)


# ---- test_analytic, quadratic BSpline net (test/test_interpolations.jl:92-102) -------------
def test_bspline_quadratic_net():
    u = np.zeros((3, 3, 3))
    u[1, 1, 1] = -3
    for I in itertools.product((0, 2), repeat=3):
        u[I] = 3
    t = np.array([-1.0, 1.0])
    f = lambda a, b, c: a * a + b * b + c * c
    for ts in ((t, t, t), (mid(t),) * 3):
        for x in itertools.product(*ts):
            got = O.point(["bspline"] * 3, [t] * 3, u, x, degrees=[2] * 3)
            assert got == pytest.approx(f(*x), abs=1e-14)


# ---- NURBS unit circle (test/test_interpolations.jl:116-149, test/test_interface.jl:270-292)
def circle_setup():
    t = np.arange(5) * (np.pi / 2)
    t_eval = np.linspace(0, 2 * np.pi, 100)
    mult = [3, 2, 2, 2, 3]
    u = np.array([[1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1], [1, -1], [1, 0]], dtype=np.float64)
    w = np.ones(9)
    w[1::2] /= np.sqrt(2)
    return t, t_eval, mult, u, w


@pytest.mark.parametrize("grid", [True, False])
def test_nurbs_unit_circle(grid):
    t, t_eval, mult, u, w = circle_setup()
    out = O.evaluate(["bspline"], [t], u, [t_eval], degrees=[2], multiplicities=[mult], weights=w, grid=grid)
    assert out.shape == (100, 2)
    np.testing.assert_allclose(out[:, 0] ** 2 + out[:, 1] ** 2, 1.0, atol=1e-14, rtol=0)
    assert len({tuple(r) for r in out[1:]}) == 99          # allunique(points[2:end])


# ---- test/test_derivatives.jl:5-79: analytic derivative == derivative of the value ---------
def _fd(f, x, h):
    # 4th-order central difference; the interpolants are polynomials per cell, h stays in-cell
    return (-f(x + 2 * h) + 8 * f(x + h) - 8 * f(x - h) + f(x - 2 * h)) / (12 * h)


@pytest.mark.parametrize("kind,deg,nu", [("linear", 0, 5), ("bspline", 2, 6), ("bspline", 5, 9)])
def test_derivative_orders_match_differentiated_value(kind, deg, nu):
    rng = np.random.default_rng(1)
    t1 = np.array([1.0, 2.0, 3.5, 4.0, 10.0])
    t2 = np.array([-1.5, 0.0, 2.5, 7.8, 12.9])
    u = rng.random((nu, nu))
    kinds, degs = [kind] * 2, [deg] * 2
    val = lambda a, b: O.point(kinds, [t1, t2], u, (a, b), degrees=degs)
    h = 1e-3
    # between data points: the value is a polynomial in a neighbourhood, so finite differences are accurate
    for a, b in itertools.product(mid(t1), mid(t2)):
        d10 = O.point(kinds, [t1, t2], u, (a, b), degrees=degs, derivative_orders=(1, 0))
        d01 = O.point(kinds, [t1, t2], u, (a, b), degrees=degs, derivative_orders=(0, 1))
        d11 = O.point(kinds, [t1, t2], u, (a, b), degrees=degs, derivative_orders=(1, 1))
        assert d10 == pytest.approx(_fd(lambda s: val(s, b), a, h), rel=1e-7, abs=1e-8)
        assert d01 == pytest.approx(_fd(lambda s: val(a, s), b, h), rel=1e-7, abs=1e-8)
        fd11 = _fd(lambda s: _fd(lambda r: val(r, s), a, h), b, h)
        assert d11 == pytest.approx(fd11, rel=1e-5, abs=1e-6)
    # in data points: the derivative is the one of the polynomial piece the index rule selects
    # (Linear: left interval, right-closed; BSpline: right interval, left-closed) == one-sided limit
    eps = 1e-9
    for a, b in itertools.product(t1, t2):
        for orders in ((1, 0), (0, 1), (1, 1)):
            at = O.point(kinds, [t1, t2], u, (a, b), degrees=degs, derivative_orders=orders)
            if kind == "linear":
                a_, b_ = (a - eps if a > t1[0] else a + eps), (b - eps if b > t2[0] else b + eps)
            else:
                a_, b_ = (a + eps if a < t1[-1] else a - eps), (b + eps if b < t2[-1] else b - eps)
            near = O.point(kinds, [t1, t2], u, (a_, b_), degrees=degs, derivative_orders=orders)
            assert at == pytest.approx(near, rel=1e-6, abs=1e-6)


# ---- test/test_datainterpolations_comparison.jl:6-75, closed forms of DataInterpolations v8 --
T1D = np.array([0.0, 0.2, 0.5, 0.7, 0.9, 1.1])
U1D = np.array([1.5, -5.3, 0.5, 3.3, 3.6, 1.0])


def _interior_points():
    return np.concatenate([np.linspace(T1D[i], T1D[i + 1], 25)[1:24] for i in range(5)])


def test_1d_linear_value_and_derivative_table():
    x = np.concatenate([_interior_points(), T1D[1:-1], T1D[[0, -1]]])
    got = O.evaluate(["linear"], [T1D], U1D, [x])
    np.testing.assert_allclose(got, np.interp(x, T1D, U1D), rtol=1e-13, atol=1e-14)
    xi = _interior_points()
    slopes = np.diff(U1D) / np.diff(T1D)
    cell = np.searchsorted(T1D, xi, side="left") - 1
    d = O.evaluate(["linear"], [T1D], U1D, [xi], derivative_orders=[1])
    np.testing.assert_allclose(d, slopes[cell], rtol=1e-13)
    # at knots: slope of the LEFT interval (first interval at t[1]); order 2 -> 0
    dk = O.evaluate(["linear"], [T1D], U1D, [T1D], derivative_orders=[1])
    np.testing.assert_allclose(dk, slopes[[0, 0, 1, 2, 3, 4]], rtol=1e-13)
    assert np.all(O.evaluate(["linear"], [T1D], U1D, [xi], derivative_orders=[2]) == 0)


def test_1d_constant_value_and_derivative_table():
    xi = _interior_points()
    got = O.evaluate(["constant"], [T1D], U1D, [xi])
    assert np.array_equal(got, U1D[np.searchsorted(T1D, xi, side="right") - 1])
    # interior knots take their own value, boundaries too (x >= t[end] -> u[end])
    assert np.array_equal(O.evaluate(["constant"], [T1D], U1D, [T1D]), U1D)
    # extrapolation: below t[1] -> u[1]; above t[end] -> u[end]
    assert np.array_equal(O.evaluate(["constant"], [T1D], U1D, [[-1.0, 7.0]]), U1D[[0, -1]])
    d = O.evaluate(["constant"], [T1D], U1D, [xi], derivative_orders=[1])
    assert np.all(d == 0)
    dk = O.evaluate(["constant"], [T1D], U1D, [T1D], derivative_orders=[1])
    assert np.all(np.isnan(dk))                      # NaN exactly on knots (:49-51)


def test_constant_derivative_leaves_vector_output_untouched():
    # src/interpolation_methods.jl:49-55 — for N_out > 0 the view is returned unchanged
    u = np.stack([U1D, 2 * U1D], axis=1)
    out = np.full((3, 2), 7.25, order="F")
    O.evaluate(["constant"], [T1D], u, [[0.1, 0.2, 0.3]], derivative_orders=[1], out=out)
    assert np.array_equal(out[[0, 2]], np.full((2, 2), 7.25))
    assert np.all(np.isnan(out[1]))


# ---- test/test_interface.jl:161-215: Float32 in -> Float32 out -----------------------------
def test_float32_support():
    t = np.array([1, 2, 3, 4], dtype=np.float32)
    u = np.random.default_rng(0).random((4, 4)).astype(np.float32)
    te = np.array([1.5, 2.5], dtype=np.float32)
    out = O.evaluate(["linear"] * 2, [t, t], u, [te, te], grid=True)
    assert out.dtype == np.float32 and out.shape == (2, 2)
    ref = O.evaluate(["linear"] * 2, [t, t], u.astype(np.float64), [te, te], grid=True)
    np.testing.assert_allclose(out, ref, rtol=1e-6)


# ---- test/test_interface.jl:217-249: unstructured multi-dimensional BSpline regression ----
def test_unstructured_multidim_bspline_regression():
    u = np.full((6, 7), 2.0)
    te = (mid(T1), mid(T2))
    out = O.evaluate(["bspline"] * 2, [T1, T2], u, te, degrees=[2, 3])
    assert out.shape == (4,)
    np.testing.assert_allclose(out, 2.0, atol=1e-13, rtol=0)
    d = O.evaluate(["bspline"] * 2, [T1, T2], u, te, degrees=[2, 3], derivative_orders=(1, 0))
    assert d.shape == (4,)
    np.testing.assert_allclose(d, 0.0, atol=1e-13, rtol=0)


# ---- basis cache (src/spline_utils.jl:164-191): partition of unity, derivative sums to 0 ----
def test_basis_cache_partition_of_unity():
    te = np.linspace(T1[0], T1[-1], 37)
    B = O.basis(T1, 3, te, max_derivative_order_eval=2)
    assert B.shape == (37, 4, 3)
    np.testing.assert_allclose(B[:, :, 0].sum(axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(B[:, :, 1].sum(axis=1), 0.0, atol=1e-13)
    assert np.all(B[:, :, 0] >= 0)


# ---- mixed kinds (NO reference semantics): slice equivalence ----------------
def test_mixed_kinds_slice_equivalence():
    rng = np.random.default_rng(3)
    tc = np.array([0.0, 1.0, 2.5, 4.0])                 # constant dimension (dim 2 of 3)
    u = rng.random((6, 4, 7))
    xa, xc, xb = rng.uniform(T1[0], T1[-1], 20), rng.uniform(-0.5, 4.5, 20), rng.uniform(T2[0], T2[-1], 20)
    got = O.evaluate(["bspline", "constant", "bspline"], [T1, tc, T2], u, [xa, xc, xb], degrees=[2, 0, 3])
    j = np.where(xc >= tc[-1], len(tc), O.get_idx("constant", tc, xc)) - 1
    for k in range(20):
        ref = O.point(["bspline"] * 2, [T1, T2], u[:, j[k], :], (xa[k], xb[k]), degrees=[2, 3])
        assert got[k] == pytest.approx(ref, rel=1e-14)
