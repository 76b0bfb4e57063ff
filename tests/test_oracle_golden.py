"""Pins the CPU oracle against the SciPy-generated golden vectors (tests/golden/make_golden.py)."""
import numpy as np

from oracle import oracle as O

RTOL = 1e-12  # norm-wise: |err| <= RTOL * max|expected| (derivatives of smooth data cancel)


def close(a, b, rtol=RTOL):
    scale = max(1.0, np.max(np.abs(b)))
    np.testing.assert_allclose(a, b, rtol=0, atol=rtol * scale)


def test_bspline_1d_all_degrees_and_derivatives(golden):
    for p in (1, 2, 3, 4, 5):
        t, u, x = (golden[f"bs1d_p{p}_{k}"] for k in "tux")
        for r in range(4):
            got = O.evaluate(["bspline"], [t], u, [x], degrees=[p], derivative_orders=[r])
            close(got, golden[f"bs1d_p{p}_r{r}"], 1e-11 if r >= 2 else RTOL)


def test_bspline_interior_multiplicities(golden):
    t, m, u, x = (golden[f"bsmult_{k}"] for k in "tmux")
    for r in (0, 1):
        got = O.evaluate(["bspline"], [t], u, [x], degrees=[2], multiplicities=[m], derivative_orders=[r])
        close(got, golden[f"bsmult_r{r}"])


def test_bspline_2d_grid_and_unstructured(golden):
    ta, tb, u, xa, xb = (golden[f"bs2d_{k}"] for k in ("ta", "tb", "u", "xa", "xb"))
    for ra, rb in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 1)):
        kw = dict(degrees=[2, 3], derivative_orders=[ra, rb])
        close(O.evaluate(["bspline"] * 2, [ta, tb], u, [xa, xb], grid=True, **kw), golden[f"bs2d_grid_r{ra}{rb}"])
        close(O.evaluate(["bspline"] * 2, [ta, tb], u, [xa, xb], grid=False, **kw), golden[f"bs2d_unst_r{ra}{rb}"])


def test_bspline_3d_cubic_gradient(golden):
    ts = [golden[f"bs3d_t{d}"] for d in range(3)]
    xs = [golden[f"bs3d_x{d}"] for d in range(3)]
    for orders in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)):
        got = O.evaluate(["bspline"] * 3, ts, golden["bs3d_u"], xs, degrees=[3] * 3, derivative_orders=orders)
        close(got, golden["bs3d_r%d%d%d" % orders])


def test_linear_nd_matches_regular_grid_interpolator(golden):
    for nd in (2, 3, 5):
        ts = [golden[f"lin{nd}d_t{d}"] for d in range(nd)]
        xs = [golden[f"lin{nd}d_x{d}"] for d in range(nd)]
        got = O.evaluate(["linear"] * nd, ts, golden[f"lin{nd}d_u"], xs)
        close(got, golden[f"lin{nd}d_val"])
