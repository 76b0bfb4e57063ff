"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every evaluation goes through the host
mirror -> C ABI (libdind.so) -> CUDA kernels and is compared with the CPU oracle on the same
seeded inputs; interval indices bit-exact, values within the north_star tolerances (helpers.TOL)."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

GENERIC = 8  # DIND_FLAG_GENERIC
CACHED_IDX = 4


@pytest.fixture(scope="module")
def D():
    import dind_b200
    assert dind_b200._lib.load().dind_device_count() > 0, "no CUDA device: the GPU tests need a B200"
    return dind_b200


DTYPES = [np.float64, np.float32]


# ---- A4/A5: interval indices, bit-exact ------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind,degree", [("linear", 0), ("constant", 0), ("bspline", 1), ("bspline", 3), ("bspline", 5)])
@pytest.mark.parametrize("uniform", [False, True])
def test_idx_eval_bit_exact(D, dtype, kind, degree, uniform):
    rng = np.random.default_rng(7)
    t = H.knots(rng, 37, dtype, uniform)
    x = H.query(rng, t, 5000, dtype)
    special = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, np.finfo(dtype).max, -np.finfo(dtype).max], dtype=dtype)
    x = np.concatenate([x, special, np.nextafter(t, dtype(np.inf)), np.nextafter(t, dtype(-np.inf))]).astype(dtype)
    dim = H.make_dims(D, [kind], [t], [x], degrees=[degree])[0]
    ref = O.get_idx(kind, t, x, degree=degree, dtype=dtype)
    assert np.array_equal(dim.idx_eval, ref)


def test_idx_eval_signed_zero_knots(D):
    for t in ([-1.0, 0.0, 1.0], [-1.0, -0.0, 1.0]):
        x = np.array([-0.0, 0.0, 1e-300, -1e-300])
        for kind in ("linear", "constant"):
            dim = H.make_dims(D, [kind], [np.array(t)], [x])[0]
            assert np.array_equal(dim.idx_eval, O.get_idx(kind, t, x))


def test_idx_eval_with_multiplicities(D):
    t = np.array([0.0, 0.7, 1.1, 2.9, 4.0])
    m = [3, 1, 2, 1, 3]
    x = np.concatenate([t, np.linspace(-1, 5, 301)])
    dim = D.BSplineInterpolationDimension(t, 2, t_eval=x, multiplicities=m)
    assert np.array_equal(dim.idx_eval, O.get_idx("bspline", t, x, degree=2, multiplicities=m))
    assert np.array_equal(dim.knots_all, O.expand_knots(t, 2, m))


# ---- A7/A8: Cox–de Boor basis cache ------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4, 5, 7])
def test_basis_function_eval_cache(D, dtype, degree):
    rng = np.random.default_rng(11)
    t = H.knots(rng, 12, dtype)
    x = H.query(rng, t, 400, dtype)
    max_r = min(degree + 1, 3)
    dim = D.BSplineInterpolationDimension(t, degree, t_eval=x, max_derivative_order_eval=max_r)
    ref = O.basis(t, degree, x, max_derivative_order_eval=max_r, dtype=dtype)
    got = dim.basis_function_eval
    assert got.shape == ref.shape == (x.size, degree + 1, max_r + 1)
    for r in range(max_r + 1):
        H.assert_close(got[:, :, r], ref[:, :, r], dtype)


# ---- A9-A13: evaluators, unstructured and grid, all kinds ------------------------------------------
def _case(rng, kinds, degrees, n_knots, out_dims, dtype, nurbs=False):
    ts = [H.knots(rng, n, dtype, uniform=(i % 2 == 1)) for i, n in enumerate(n_knots)]
    n_u = []
    for k, p, t in zip(kinds, degrees, ts):
        n_u.append(t.size + p - 1 if k == "bspline" else t.size)
    u = rng.random(tuple(n_u) + tuple(out_dims)).astype(dtype)
    w = (0.5 + rng.random(tuple(n_u))).astype(dtype) if nurbs else None
    return ts, u, w


CASES = [
    # kinds, degrees, knots per dim, out dims, derivative orders to test
    (("linear",), (0,), (9,), (), [(0,), (1,), (2,)]),
    (("linear",) * 2, (0, 0), (8, 7), (2,), [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0)]),
    (("linear",) * 3, (0,) * 3, (6, 5, 7), (), [(0, 0, 0), (0, 1, 0)]),
    (("linear",) * 5, (0,) * 5, (4, 3, 4, 3, 5), (4,), [(0,) * 5, (0, 0, 1, 0, 0)]),
    (("constant",), (0,), (9,), (), [(0,), (1,)]),
    (("constant",) * 2, (0, 0), (6, 7), (3,), [(0, 0), (1, 0)]),
    (("bspline",), (3,), (9,), (), [(0,), (1,), (2,), (3,)]),
    (("bspline",) * 2, (2, 3), (7, 6), (2,), [(0, 0), (1, 0), (0, 1), (1, 1), (2, 1)]),
    (("bspline",) * 2, (5, 1), (7, 6), (), [(0, 0), (1, 1)]),
    (("bspline",) * 3, (3, 3, 3), (6, 5, 7), (3,), [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]),
    (("bspline",) * 4, (2,) * 4, (5, 4, 5, 4), (), [(0,) * 4]),
    (("bspline", "constant", "bspline"), (2, 0, 3), (6, 5, 6), (), [(0, 0, 0), (1, 0, 0)]),
    (("linear", "bspline"), (0, 2), (6, 5), (2,), [(0, 0), (1, 1)]),
    (("constant", "linear", "bspline"), (0, 0, 1), (4, 5, 6), (), [(0, 0, 0), (0, 1, 0)]),
    (("bspline", "bspline", "constant", "bspline"), (2, 2, 0, 2), (5, 4, 5, 4), (), [(0, 0, 0, 0), (1, 0, 0, 0)]),
    (("linear", "linear", "constant", "constant", "linear"), (0,) * 5, (4, 3, 4, 3, 4), (2,), [(0,) * 5]),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("flags", [0, GENERIC])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_unstructured_parity(D, dtype, flags, case):
    kinds, degrees, n_knots, out_dims, orders_list = CASES[case]
    rng = np.random.default_rng(100 + case)
    ts, u, _ = _case(rng, kinds, degrees, n_knots, out_dims, dtype)
    xs = H.zipped_queries(rng, ts, 3000, dtype)
    for orders in orders_list:
        kw = dict(degrees=degrees, derivative_orders=orders)
        ref = H.oracle_eval(kinds, ts, u, xs, **kw)
        if "constant" in kinds and any(o > 0 for o in orders) and out_dims:
            out = np.full(ref.shape, 7.25, dtype=dtype, order="F")
            ref = O.evaluate(kinds, ts, u, xs, out=out.copy(order="F"), **kw)
            got, _ = H.gpu_eval(D, kinds, ts, u, xs, flags=flags, out=out, **kw)
        else:
            got, _ = H.gpu_eval(D, kinds, ts, u, xs, flags=flags, **kw)
        H.assert_close(got, ref, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("flags", [0, GENERIC])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_grid_parity(D, dtype, flags, case):
    kinds, degrees, n_knots, out_dims, orders_list = CASES[case]
    rng = np.random.default_rng(200 + case)
    ts, u, _ = _case(rng, kinds, degrees, n_knots, out_dims, dtype)
    n_per = {1: 600, 2: 45, 3: 14, 4: 7, 5: 5}[len(kinds)]
    xs = [np.sort(H.query(rng, t, n_per, dtype)) if i % 2 == 0 else H.query(rng, t, n_per, dtype) for i, t in enumerate(ts)]
    for orders in orders_list:
        kw = dict(degrees=degrees, derivative_orders=orders, grid=True)
        ref = H.oracle_eval(kinds, ts, u, xs, **kw)
        if "constant" in kinds and any(o > 0 for o in orders) and out_dims:
            out = np.full(ref.shape, 7.25, dtype=dtype, order="F")
            ref = O.evaluate(kinds, ts, u, xs, out=out.copy(order="F"), **kw)
            got, _ = H.gpu_eval(D, kinds, ts, u, xs, flags=flags, out=out, **kw)
        else:
            got, _ = H.gpu_eval(D, kinds, ts, u, xs, flags=flags, **kw)
        H.assert_close(got, ref, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("flags", [0, GENERIC])
@pytest.mark.parametrize("n_in,degree,out_dims", [(1, 2, (2,)), (2, 3, ()), (2, 1, (3,)), (3, 2, ()), (4, 2, ())])
def test_nurbs_parity(D, dtype, flags, n_in, degree, out_dims):
    rng = np.random.default_rng(300 + n_in)
    kinds, degrees = ("bspline",) * n_in, (degree,) * n_in
    ts, u, w = _case(rng, kinds, degrees, (6, 5, 6, 5)[:n_in], out_dims, dtype, nurbs=True)
    # inside the knot span (norm-wise bound); outside it the quotient has poles and
    # test_nurbs_extrapolation_parity applies the point-wise quotient bound instead
    xs = H.zipped_queries(rng, ts, 2000, dtype, outside=False)
    ref = H.oracle_eval(kinds, ts, u, xs, degrees=degrees, weights=w)
    got, _ = H.gpu_eval(D, kinds, ts, u, xs, degrees=degrees, weights=w, flags=flags)
    H.assert_close(got, ref, dtype)
    n_per = {1: 300, 2: 40, 3: 12, 4: 6}[n_in]
    xg = [np.sort(H.query(rng, t, n_per, dtype, outside=False)) for t in ts]
    ref = H.oracle_eval(kinds, ts, u, xg, degrees=degrees, weights=w, grid=True)
    got, _ = H.gpu_eval(D, kinds, ts, u, xg, degrees=degrees, weights=w, grid=True, flags=flags)
    H.assert_close(got, ref, dtype)


def test_eval_unstructured_allgather_without_a_communicator(D):
    """dind_eval_unstructured_allgather on ONE process (world = 1): the shard is the whole point list, the call is
    eval_unstructured into the global array — in any number of chunks, bit for bit — and it validates its arguments
    (the multi-GPU behaviour is tested by tests/c_abi_allgather.c and tests/multi_gpu_worker.py on >= 2 devices)."""
    import torch
    rng = np.random.default_rng(99)
    kinds, degrees = ("linear",) * 3, (0,) * 3
    ts, u, _ = _case(rng, kinds, degrees, (6, 5, 7), (3,), np.float64)
    xs = H.zipped_queries(rng, ts, 30001, np.float64)
    n = xs[0].size
    itp = D.NDInterpolation(u, H.make_dims(D, kinds, ts, [torch.from_numpy(x).cuda() for x in xs], degrees))
    ref = D.eval_unstructured(itp, derivative_orders=(0, 1, 0))
    ref = ref.cpu().numpy() if hasattr(ref, "cpu") else np.asarray(ref)
    for chunks in (0, 1, 7, 64, 1000):
        full = D.f_empty((n, 3), np.float64, "cuda")
        full.fill_(float("nan"))
        D.eval_unstructured_allgather_(full, itp, n, derivative_orders=(0, 1, 0), n_chunks=chunks)
        torch.cuda.synchronize()
        assert np.array_equal(full.cpu().numpy(), ref), chunks
    with pytest.raises(AssertionError):                       # the handle holds n points, not n + 1
        D.eval_unstructured_allgather_(D.f_empty((n + 1, 3), np.float64, "cuda"), itp, n + 1)
    with pytest.raises(AssertionError):                       # host array
        D.eval_unstructured_allgather_(np.zeros((n, 3), order="F"), itp, n)
    with pytest.raises(AssertionError):                       # wrong shape
        D.eval_unstructured_allgather_(D.f_empty((n, 2), np.float64, "cuda"), itp, n)
    itp.close()


@pytest.mark.parametrize("out_dims", [(), (2,)])
def test_nurbs_grid_denominator_cache_follows_weights_knots_and_t_eval(D, out_dims, monkeypatch):
    """4-D NURBS grids (two-stage path): the stage-1 result of the WEIGHTS does not depend on u and is re-used when
    eval_grid! is repeated with a new u (BASELINE config 4 in a loop).  It must be recomputed after new weights, new
    t_eval, and never leak between them: every step equals a fresh handle with DIND_NO_DEN_CACHE bit for bit, and the
    oracle within the tolerance."""
    rng = np.random.default_rng(4242)
    kinds, degrees = ("bspline",) * 4, (2,) * 4
    ts, u, w = _case(rng, kinds, degrees, (7, 6, 7, 6), out_dims, np.float64, nurbs=True)
    mkx = lambda n: [np.sort(H.query(rng, t, m, np.float64, outside=False)) for t, m in zip(ts, n)]
    xg = mkx((12, 10, 11, 9))

    def fresh(u_, w_, x_):
        monkeypatch.setenv("DIND_NO_DEN_CACHE", "1")
        itp = D.NDInterpolation(u_, H.make_dims(D, kinds, ts, x_, degrees), cache=D.NURBSWeights(w_))
        monkeypatch.delenv("DIND_NO_DEN_CACHE")
        out = D.eval_grid(itp)
        assert itp.last_kernel.startswith("grid_split"), itp.last_kernel
        itp.close()
        return out

    itp = D.NDInterpolation(u, H.make_dims(D, kinds, ts, xg, degrees), cache=D.NURBSWeights(w))
    steps = []
    u1 = rng.random(u.shape)
    w1 = 0.5 + rng.random(w.shape)
    xg1 = mkx((12, 10, 11, 9))
    xg2 = mkx((10, 12, 9, 13))
    steps.append(("first", lambda: None, u, w, xg))
    steps.append(("new u", lambda: itp.set_u(u1), u1, w, xg))
    steps.append(("same u again", lambda: None, u1, w, xg))
    steps.append(("new weights", lambda: itp._set_weights(w1), u1, w1, xg))
    steps.append(("new u after new weights", lambda: itp.set_u(u), u, w1, xg))
    steps.append(("new t_eval, same sizes", lambda: itp.set_t_eval(xg1), u, w1, xg1))
    steps.append(("new t_eval, other sizes", lambda: itp.set_t_eval(xg2), u, w1, xg2))
    steps.append(("new u at the end", lambda: itp.set_u(u1), u1, w1, xg2))
    for what, act, u_, w_, x_ in steps:
        act()
        got = D.eval_grid(itp)
        assert itp.last_kernel.startswith("grid_split"), itp.last_kernel
        assert np.array_equal(got, fresh(u_, w_, x_)), what
        H.assert_close(got, H.oracle_eval(kinds, ts, u_, x_, degrees=degrees, weights=w_, grid=True), np.float64)
    itp.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("flags", [0, GENERIC])
@pytest.mark.parametrize("grid", [False, True])
@pytest.mark.parametrize("n_in,degree,out_dims", [(1, 2, (2,)), (2, 3, ()), (3, 2, ()), (4, 2, ())])
def test_nurbs_extrapolation_parity(D, dtype, flags, grid, n_in, degree, out_dims):
    """NURBS OUTSIDE the knot span: the reference evaluates there with the clamped interval
    (src/interpolation_utils.jl:128-132) and divides (src/interpolation_methods.jl:134-138).  The quotient of two
    extrapolated polynomials has poles, so the bound is the first-order one for a quotient whose numerator
    sum(w*B*u) and denominator sum(w*B) are each accurate to TOL norm-wise:
        |q_gpu - q_ref| * |den_ref| <= TOL * (max|num_ref| + |q_ref| * max|den_ref|)      (point-wise)
    on every point whose denominator is at least 5 % of the typical one — and those must be > 80 % of the points.
    Numerator and denominator are also checked on their own: both are plain B-spline contractions (of u*w and of
    w), evaluated by the same library outside the span and compared with the oracle norm-wise at TOL."""
    rng = np.random.default_rng(700 + 10 * n_in + degree)
    kinds, degrees = ("bspline",) * n_in, (degree,) * n_in
    ts, u, w = _case(rng, kinds, degrees, (6, 5, 6, 5)[:n_in], out_dims, dtype, nurbs=True)
    def pts(t, n):   # 60 % inside the span, 40 % up to 10 % of the span outside it, plus both ends of the span
        span = float(t[-1] - t[0])
        k = int(0.4 * n)
        side = rng.random(k) < 0.5
        out = np.where(side, t[0] - rng.random(k) * 0.1 * span, t[-1] + rng.random(k) * 0.1 * span)
        x = np.concatenate([rng.uniform(t[0], t[-1], n - k - 2), out, [t[0], t[-1]]])
        rng.shuffle(x)
        return x.astype(dtype)
    if grid:
        n_per = {1: 300, 2: 40, 3: 12, 4: 6}[n_in]
        xs = [np.sort(pts(t, n_per)) for t in ts]
    else:
        xs = [pts(t, 2000) for t in ts]
    kw = dict(degrees=degrees, grid=grid)
    wshape = w.shape + (1,) * len(out_dims)
    uw = (u * w.reshape(wshape)).astype(dtype)
    num_ref = H.oracle_eval(kinds, ts, uw, xs, **kw)
    den_ref = H.oracle_eval(kinds, ts, w, xs, **kw)
    q_ref = H.oracle_eval(kinds, ts, u, xs, weights=w, **kw)
    num_gpu, _ = H.gpu_eval(D, kinds, ts, uw, xs, flags=flags, **kw)
    den_gpu, _ = H.gpu_eval(D, kinds, ts, w, xs, flags=flags, **kw)
    H.assert_close(num_gpu, num_ref, dtype)
    H.assert_close(den_gpu, den_ref, dtype)
    q_gpu, _ = H.gpu_eval(D, kinds, ts, u, xs, weights=w, flags=flags, **kw)
    tol = H.TOL[np.dtype(dtype)]
    den_b = den_ref.reshape(den_ref.shape + (1,) * len(out_dims)).astype(np.float64)
    s_num, s_den = float(np.max(np.abs(num_ref))), float(np.max(np.abs(den_ref)))
    ok = np.abs(den_b) >= 0.05 * float(np.median(np.abs(den_ref)))
    ok = np.broadcast_to(ok, q_ref.shape)
    assert ok.mean() > 0.8, ok.mean()
    lhs = np.abs(q_gpu.astype(np.float64) - q_ref.astype(np.float64)) * np.abs(den_b)
    rhs = tol * (s_num + np.abs(q_ref.astype(np.float64)) * s_den)
    H.record_observed(float(np.max((lhs / rhs)[ok])), float(np.max((lhs / rhs)[ok])) * tol)
    assert np.all(lhs[ok] <= rhs[ok]), float(np.max((lhs / rhs)[ok]))
    assert np.array_equal(np.isnan(q_gpu), np.isnan(q_ref))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("degree,nurbs,out_dims,orders", [
    (2, True, (), (0, 0, 0, 0)), (2, False, (2,), (0, 1, 0, 1)), (3, True, (2,), (0, 0, 0, 0)), (3, False, (), (1, 0, 2, 0))])
@pytest.mark.parametrize("layout", ["upsample", "ragged", "unsorted"])
def test_grid_split_4d_matches_oracle_and_pencil(D, dtype, degree, nurbs, out_dims, orders, layout):
    """4-D B-spline / NURBS grids go through the two-stage separable path (dind_fast_grid_split.cu) when the
    grid is at least as fine as the nodes along axes 3 and 4: same numbers as the oracle and as the
    one-pass kernels (which DIND_FLAG_GENERIC forces)."""
    rng = np.random.default_rng(900 + degree)
    kinds, degrees = ("bspline",) * 4, (degree,) * 4
    ts, u, w = _case(rng, kinds, degrees, (5, 4, 6, 5), out_dims, dtype, nurbs=nurbs)
    n_g = {"upsample": (70, 9, 14, 12), "ragged": (33, 3, 9, 11), "unsorted": (8, 5, 16, 13)}[layout]
    xg = []
    for t, n in zip(ts, n_g):
        x = rng.uniform(t[0], t[-1], n).astype(dtype)
        x[0] = t[0]; x[-1] = t[-1]                    # both ends of the knot span are grid points
        xg.append(x if layout == "unsorted" else np.sort(x))
    kw = dict(degrees=degrees, weights=w, grid=True, derivative_orders=orders)
    ref = H.oracle_eval(kinds, ts, u, xg, **kw)
    got, kernel = H.gpu_eval(D, kinds, ts, u, xg, **kw)
    assert kernel.startswith("grid_split"), kernel
    H.assert_close(got, ref, dtype)
    gen, kernel_g = H.gpu_eval(D, kinds, ts, u, xg, flags=GENERIC, **kw)
    assert kernel_g == "generic_grid"
    H.assert_close(got, gen, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("const_axis", [2, 3])
@pytest.mark.parametrize("degree,nurbs,out_dims", [(2, True, ()), (3, False, (2,)), (2, False, ())])
def test_grid_split_4d_with_constant_selector_axis(D, dtype, const_axis, degree, nurbs, out_dims):
    """BASELINE config 4's mixed variant (a Constant dimension next to B-spline ones) through the two-stage
    path: the Constant axis only selects a slice in stage 2.  Pinned by the oracle's tensor-product semantics
    and by slice equivalence with the pure B-spline evaluation on the selected slices."""
    rng = np.random.default_rng(950 + degree + const_axis)
    kinds = tuple("constant" if d == const_axis else "bspline" for d in range(4))
    degrees = tuple(0 if d == const_axis else degree for d in range(4))
    ts, u, w = _case(rng, kinds, degrees, (5, 4, 6, 5), out_dims, dtype, nurbs=nurbs)
    xg = []
    for d, (t, n) in enumerate(zip(ts, (40, 7, 13, 12))):
        x = np.sort(rng.uniform(t[0], t[-1], n).astype(dtype))
        if d == const_axis:
            x[:3] = t[:3]                    # exactly on knots, and beyond the last knot
            x[-1] = t[-1] + dtype(0.25)
            x = np.sort(x)
        xg.append(x)
    kw = dict(degrees=degrees, weights=w, grid=True)
    ref = H.oracle_eval(kinds, ts, u, xg, **kw)
    got, kernel = H.gpu_eval(D, kinds, ts, u, xg, **kw)
    assert kernel.startswith("grid_split") and "selector" in kernel, kernel
    H.assert_close(got, ref, dtype)
    # slice equivalence: fix the Constant axis at the node the reference's index rule picks
    tc = ts[const_axis]
    j = np.clip(np.searchsorted(tc, xg[const_axis], side="right"), 1, tc.size - 1) - 1
    j = np.where(xg[const_axis] >= tc[-1], tc.size - 1, j)
    rest = [d for d in range(4) if d != const_axis]
    for gi in (0, 2, len(j) - 1):
        us = np.take(u, j[gi], axis=const_axis)
        ws = None if w is None else np.take(w, j[gi], axis=const_axis)
        sl = H.oracle_eval(("bspline",) * 3, [ts[d] for d in rest], us, [xg[d] for d in rest],
                           degrees=(degree,) * 3, weights=ws, grid=True)
        H.assert_close(np.take(got, gi, axis=const_axis), sl, dtype)


def test_cached_idx_path_matches(D):
    rng = np.random.default_rng(5)
    kinds, degrees = ("bspline",) * 2, (3, 2)
    ts, u, _ = _case(rng, kinds, degrees, (7, 6), (2,), np.float64)
    xs = H.zipped_queries(rng, ts, 2000, np.float64)
    a, _ = H.gpu_eval(D, kinds, ts, u, xs, degrees=degrees, flags=0)
    b, _ = H.gpu_eval(D, kinds, ts, u, xs, degrees=degrees, flags=CACHED_IDX)
    c, _ = H.gpu_eval(D, kinds, ts, u, xs, degrees=degrees, flags=CACHED_IDX | GENERIC)
    assert np.array_equal(a, b)
    H.assert_close(c, a, np.float64)


# ---- the reference's known-answer tests, on the GPU --------------------------------------------
T1 = np.array([-3.14, 1.0, 3.0, 7.6, 12.8])
T2 = np.array([-2.71, 1.41, 12.76, 50.2, 120.0])


def _mid(t):
    return t[:-1] + np.diff(t) / 2


@pytest.mark.parametrize("case", ["linear", "bspline", "nurbs"])
def test_kat_globally_constant(D, case):   # test/test_interpolations.jl:5-50, :68, :86-90, :108-114
    rng = np.random.default_rng(10)
    if case == "linear":
        kinds, degs, u, w, derivs = ["linear"] * 2, None, np.full((5, 5), 2.0), None, True
    elif case == "bspline":
        kinds, degs, u, w, derivs = ["bspline"] * 2, [2, 3], np.full((6, 7), 2.0), None, True
    else:
        kinds, degs, u, w, derivs = ["bspline"] * 2, [3, 1], np.full((7, 5), 2.0), rng.random((7, 5)), False
    for te in ((T1, T2), (_mid(T1), _mid(T2))):
        val, _ = H.gpu_eval(D, kinds, [T1, T2], u, te, degrees=degs, weights=w, grid=True)
        np.testing.assert_allclose(val, 2.0, atol=1e-10, rtol=0)
        if derivs:
            for orders in ((1, 0), (0, 1)):
                d, _ = H.gpu_eval(D, kinds, [T1, T2], u, te, degrees=degs, derivative_orders=orders, grid=True)
                np.testing.assert_allclose(d, 0.0, atol=1e-10, rtol=0)


def test_kat_linear_plane_single_point(D):   # test/test_interpolations.jl:52-82
    rng = np.random.default_rng(1)
    t1, t2 = np.cumsum(rng.random(10)), np.cumsum(rng.random(10))
    f = lambda a, b: 3.0 + 2.3 * a - 4.7 * b
    itp = D.NDInterpolation(f(t1[:, None], t2[None, :]), (D.LinearInterpolationDimension(t1), D.LinearInterpolationDimension(t2)))
    for ts in ((t1, t2), (_mid(t1), _mid(t2))):
        for a, b in itertools.product(*ts):
            assert itp(a, b) == pytest.approx(f(a, b), rel=1e-12, abs=1e-12)
    itp.close()


def test_kat_bspline_quadratic_net(D):   # test/test_interpolations.jl:92-102
    u = np.zeros((3, 3, 3))
    u[1, 1, 1] = -3
    for I in itertools.product((0, 2), repeat=3):
        u[I] = 3
    t = np.array([-1.0, 1.0])
    dim = D.BSplineInterpolationDimension(t, 2)
    itp = D.NDInterpolation(u, (dim, dim, dim))
    for ts in ((t,) * 3, (_mid(t),) * 3):
        for x in itertools.product(*ts):
            assert itp(*x) == pytest.approx(sum(v * v for v in x), abs=1e-13)
    itp.close()


@pytest.mark.parametrize("grid", [True, False])
def test_kat_nurbs_unit_circle(D, grid):   # test/test_interpolations.jl:116-149, test/test_interface.jl:270-292
    t = np.arange(5) * (np.pi / 2)
    t_eval = np.linspace(0, 2 * np.pi, 100)
    u = np.array([[1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1], [1, -1], [1, 0]], dtype=np.float64)
    w = np.ones(9)
    w[1::2] /= np.sqrt(2)
    dim = D.BSplineInterpolationDimension(t, 2, multiplicities=[3, 2, 2, 2, 3], t_eval=t_eval)
    itp = D.NDInterpolation(u, dim, cache=D.NURBSWeights(w))
    out = (D.eval_grid if grid else D.eval_unstructured)(itp)
    assert out.shape == (100, 2)
    np.testing.assert_allclose(out[:, 0] ** 2 + out[:, 1] ** 2, 1.0, atol=1e-13, rtol=0)
    assert len({tuple(r) for r in out[1:]}) == 99
    itp.close()


def test_kat_1d_tables(D):   # test/test_datainterpolations_comparison.jl:6-75
    t = np.array([0.0, 0.2, 0.5, 0.7, 0.9, 1.1])
    u = np.array([1.5, -5.3, 0.5, 3.3, 3.6, 1.0])
    xi = np.concatenate([np.linspace(t[i], t[i + 1], 25)[1:24] for i in range(5)])
    x = np.concatenate([xi, t])
    lin = D.NDInterpolation(u, D.LinearInterpolationDimension(t, t_eval=x))
    np.testing.assert_allclose(D.eval_unstructured(lin), np.interp(x, t, u), rtol=1e-12, atol=1e-13)
    slopes = np.diff(u) / np.diff(t)
    d = D.eval_unstructured(lin, derivative_orders=(1,))
    np.testing.assert_allclose(d[: xi.size], slopes[np.searchsorted(t, xi, side="left") - 1], rtol=1e-12)
    np.testing.assert_allclose(d[xi.size:], slopes[[0, 0, 1, 2, 3, 4]], rtol=1e-12)
    con = D.NDInterpolation(u, D.ConstantInterpolationDimension(t, t_eval=x))
    v = D.eval_unstructured(con)
    assert np.array_equal(v[: xi.size], u[np.searchsorted(t, xi, side="right") - 1])
    assert np.array_equal(v[xi.size:], u)
    dc = D.eval_unstructured(con, derivative_orders=(1,))
    assert np.all(dc[: xi.size] == 0) and np.all(np.isnan(dc[xi.size:]))
    lin.close(); con.close()


def test_golden_scipy_vectors_on_gpu(D, golden):
    for p in (1, 2, 3, 4, 5):
        t, u, x = (golden[f"bs1d_p{p}_{k}"] for k in "tux")
        for r in range(4):
            got, _ = H.gpu_eval(D, ["bspline"], [t], u, [x], degrees=[p], derivative_orders=[r])
            H.assert_close(got, golden[f"bs1d_p{p}_r{r}"], np.float64)
    ta, tb, u, xa, xb = (golden[f"bs2d_{k}"] for k in ("ta", "tb", "u", "xa", "xb"))
    for ra, rb in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 1)):
        kw = dict(degrees=[2, 3], derivative_orders=[ra, rb])
        got, _ = H.gpu_eval(D, ["bspline"] * 2, [ta, tb], u, [xa, xb], grid=True, **kw)
        H.assert_close(got, golden[f"bs2d_grid_r{ra}{rb}"], np.float64)
        got, _ = H.gpu_eval(D, ["bspline"] * 2, [ta, tb], u, [xa, xb], **kw)
        H.assert_close(got, golden[f"bs2d_unst_r{ra}{rb}"], np.float64)
    for nd in (2, 3, 5):
        ts = [golden[f"lin{nd}d_t{d}"] for d in range(nd)]
        xs = [golden[f"lin{nd}d_x{d}"] for d in range(nd)]
        got, _ = H.gpu_eval(D, ["linear"] * nd, ts, golden[f"lin{nd}d_u"], xs)
        H.assert_close(got, golden[f"lin{nd}d_val"], np.float64)


# ---- interface behaviour ---------------------------------------------------------------------------
def test_float32_in_float32_out_and_promotion(D):   # test/test_interface.jl:161-215
    t = np.array([1, 2, 3, 4], dtype=np.float32)
    u = np.random.default_rng(0).random((4, 4)).astype(np.float32)
    te = np.array([1.5, 2.5], dtype=np.float32)
    dims = (D.LinearInterpolationDimension(t, t_eval=te),) * 2
    itp = D.NDInterpolation(u, dims)
    out = D.eval_grid(itp)
    assert out.dtype == np.float32 and out.shape == (2, 2)
    assert itp.dtype == np.float32
    itp64 = D.NDInterpolation(u.astype(np.float64), (D.LinearInterpolationDimension(t),) * 2)
    assert itp64.dtype == np.float64
    assert isinstance(float(itp64(1.5, 2.5)), float)
    itp.close(); itp64.close()


def test_validation_errors(D):
    t = np.array([0.0, 1.0, 2.0])
    with pytest.raises(AssertionError):
        D.LinearInterpolationDimension([0.0, 0.0, 1.0])
    with pytest.raises(AssertionError):   # validate_size_u
        D.NDInterpolation(np.zeros((4, 3)), (D.LinearInterpolationDimension(t),) * 2)
    with pytest.raises(AssertionError):   # multiplicities out of range
        D.BSplineInterpolationDimension(t, 2, multiplicities=[3, 4, 3])
    bs = D.BSplineInterpolationDimension(t, 2, t_eval=[0.5, 1.5])
    itp = D.NDInterpolation(np.zeros(4), bs)
    with pytest.raises(AssertionError):   # derivative order > max_derivative_order_eval
        D.eval_unstructured(itp, derivative_orders=(1,))
    with pytest.raises(AssertionError):   # negative order
        D.eval_unstructured(itp, derivative_orders=(-1,))
    itp.close()
    nur = D.NDInterpolation(np.zeros(4), D.BSplineInterpolationDimension(t, 2, t_eval=[0.5], max_derivative_order_eval=1),
                            cache=D.NURBSWeights(np.ones(4)))
    with pytest.raises(AssertionError):   # NURBS derivatives unsupported
        D.eval_unstructured(nur, derivative_orders=(1,))
    nur.close()
    with pytest.raises(AssertionError):   # weights size
        D.NDInterpolation(np.zeros(4), D.BSplineInterpolationDimension(t, 2), cache=D.NURBSWeights(np.ones(5)))
    itp2 = D.NDInterpolation(np.zeros((3, 3)), (D.LinearInterpolationDimension(t, t_eval=[0.1, 0.2]),
                                                D.LinearInterpolationDimension(t, t_eval=[0.1])))
    with pytest.raises(AssertionError):   # unequal t_eval lengths for unstructured
        D.eval_unstructured(itp2)
    assert D.eval_grid(itp2).shape == (2, 1)
    itp2.close()


CELL_SORTED = 16


def test_single_point_in_place_and_eval_points(D):   # src/DataInterpolationsND.jl:85-94
    rng = np.random.default_rng(2)
    t1, t2 = H.knots(rng, 6), H.knots(rng, 7)
    u = rng.random((6, 7, 3))
    itp = D.NDInterpolation(u, (D.LinearInterpolationDimension(t1), D.LinearInterpolationDimension(t2)))
    out = np.zeros(3)
    itp(2.0, 3.0, out=out)
    ref = O.point(["linear"] * 2, [t1, t2], u, (2.0, 3.0))
    np.testing.assert_allclose(out, ref, rtol=1e-13)
    with pytest.raises(AssertionError):
        itp(2.0, 3.0, out=np.zeros(4))
    xs = [rng.uniform(t1[0], t1[-1], 100), rng.uniform(t2[0], t2[-1], 100)]
    got = itp.eval_points(xs, derivative_orders=(1, 0))
    H.assert_close(got, O.evaluate(["linear"] * 2, [t1, t2], u, xs, derivative_orders=(1, 0)), np.float64)
    itp.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 31, 2048, 2049, 5000])
def test_eval_points_small_batch_path_equals_copy_path(D, dtype, n):
    """dind_eval_points takes a pinned, device-mapped staging path for small all-host batches
    (n <= 2048); results must be bit-identical to the copy path and to the cached evaluation."""
    rng = np.random.default_rng(40 + n)
    ts = [H.knots(rng, 9, dtype), H.knots(rng, 8, dtype), H.knots(rng, 7, dtype)]
    u = rng.random((10, 9, 8, 2)).astype(dtype)
    xs = [rng.uniform(t[0], t[-1], n).astype(dtype) for t in ts]
    dims = tuple(D.BSplineInterpolationDimension(t, 2, t_eval=x, max_derivative_order_eval=1) for t, x in zip(ts, xs))
    itp = D.NDInterpolation(u, dims)
    for orders in [(0, 0, 0), (1, 0, 1)]:
        got = itp.eval_points(xs, derivative_orders=orders)                      # small path for n <= 2048
        cached = D.eval_unstructured(itp, derivative_orders=orders)
        np.testing.assert_array_equal(got, cached)
        import torch
        dev = itp.eval_points([torch.from_numpy(x).cuda() for x in xs], derivative_orders=orders)   # device pointers: no staging
        np.testing.assert_array_equal(dev.cpu().numpy(), cached)
        ref = O.evaluate(["bspline"] * 3, ts, u, xs, degrees=[2] * 3, derivative_orders=orders)
        H.assert_close(got, ref, dtype)
    itp.close()


def test_eval_points_small_batch_keeps_untouched_semantics(D):   # src/interpolation_methods.jl:49-55
    t = np.array([0.0, 1.0, 2.0, 3.0])
    u = np.arange(8.0).reshape(4, 2)
    itp = D.NDInterpolation(u, D.ConstantInterpolationDimension(t))
    out = np.full((3, 2), 7.0, order="F")
    itp.eval_points([np.array([0.5, 1.0, 2.5])], out=out, derivative_orders=(1,))
    assert np.all(out[[0, 2]] == 7.0) and np.all(np.isnan(out[1]))     # off the knots: untouched; on a knot: NaN
    one = np.full(2, 5.0)
    itp(0.5, out=one, derivative_orders=(1,))
    assert np.all(one == 5.0)
    itp.close()


def test_build_caches_is_optional_and_idempotent(D):
    rng = np.random.default_rng(77)
    ts = [H.knots(rng, 12), H.knots(rng, 10)]
    u = rng.random((12, 10, 3))
    xs = [rng.uniform(t[0] - 0.3, t[-1] + 0.3, 4000) for t in ts]
    mk = lambda: D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs)))
    a, b = mk(), mk()
    n0 = b.launch_count
    b.build_caches(flags=CELL_SORTED | 4)          # plan + idx cache + component-fastest u, no evaluation
    n1 = b.launch_count
    assert n1 > n0
    b.build_caches(flags=CELL_SORTED | 4)          # nothing left to build
    assert b.launch_count == n1
    assert b.plan_permutation().shape == (4000,)
    np.testing.assert_array_equal(D.eval_unstructured(b, flags=CELL_SORTED), D.eval_unstructured(a))
    assert b.launch_count - n1 == 2                # evaluation kernel + un-permutation only
    g = D.NDInterpolation(rng.random((12, 10)), tuple(D.LinearInterpolationDimension(t, t_eval=np.sort(x[:50])) for t, x in zip(ts, xs)))
    g.build_caches(grid=True, derivative_orders=(1, 0))
    n2 = g.launch_count
    D.eval_grid(g, derivative_orders=(1, 0))
    assert g.launch_count - n2 == 1                # tables were ready
    with pytest.raises(AssertionError):
        D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t) for t in ts)).build_caches(derivative_orders=(0, -1))
    for i in (a, b, g):
        i.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2047, 2048, 2049, 4097, 70001])
@pytest.mark.parametrize("comps", [(), (2,), (3,), (4,), (5,)])
@pytest.mark.parametrize("scheme", ["by-size", "buckets", "original-order", "buckets, per-dimension gathers"])
def test_cell_sorted_window_boundaries_and_record_shapes(D, dtype, n, comps, scheme, monkeypatch):
    """The plan brings results home through staging records of 8 / 16 / 32 / n*16 bytes, either in slots ordered by
    windows of 2048 original points (records smaller than a sector) or at their original index (DIND_PLAN_BUCKETS
    forces either): every window fill level and every record shape must give the unsorted result bit for bit (value
    and fused gradient), and a re-plan after new t_eval must too.  DIND_NO_PLAN_RECS: sorted t_eval copies by one
    gather per dimension instead of one record read per point."""
    if scheme != "by-size":
        monkeypatch.setenv("DIND_PLAN_BUCKETS", "0" if scheme == "original-order" else "1")
    if scheme.endswith("gathers"):
        monkeypatch.setenv("DIND_NO_PLAN_RECS", "1")
    rng = np.random.default_rng(n + 17 * len(comps) + (comps[0] if comps else 0))
    ts = [H.knots(rng, 7, dtype), H.knots(rng, 6, dtype)]
    u = rng.random((7, 6) + comps).astype(dtype)
    mk = lambda xs: D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs)))
    for rep in range(2):
        xs = [rng.uniform(t[0] - 0.2, t[-1] + 0.2, n).astype(dtype) for t in ts]
        if rep == 0:
            a, b = mk(xs), mk(xs)
        else:
            a.set_t_eval(xs); b.set_t_eval(xs)          # re-plan on the same handle
        for orders in [(0, 0), (1, 0)]:
            np.testing.assert_array_equal(D.eval_unstructured(b, derivative_orders=orders, flags=CELL_SORTED),
                                          D.eval_unstructured(a, derivative_orders=orders))
        np.testing.assert_array_equal(D.eval_unstructured_gradient(b, flags=CELL_SORTED), D.eval_unstructured_gradient(a))
        perm = b.plan_permutation()
        assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    a.close(); b.close()


def test_empty_and_single_point_batches(D):
    t = np.array([0.0, 1.0, 2.0])
    itp = D.NDInterpolation(np.array([1.0, 2.0, 4.0]), D.LinearInterpolationDimension(t))
    assert D.eval_unstructured(itp).shape == (0,)          # t_eval defaults to no points
    assert D.eval_grid(itp).shape == (0,)
    itp.close()
    itp = D.NDInterpolation(np.array([1.0, 2.0, 4.0]), D.LinearInterpolationDimension(t, t_eval=[1.5]))
    assert D.eval_unstructured(itp).tolist() == [3.0]
    itp.close()


def test_device_tensors_zero_copy_and_new_u(D):
    import torch
    rng = np.random.default_rng(4)
    ts = [H.knots(rng, 9, np.float32) for _ in range(3)]
    xs = [np.sort(rng.uniform(t[0], t[-1], 17)).astype(np.float32) for t in ts]
    dims = tuple(D.LinearInterpolationDimension(t, t_eval=torch.from_numpy(x).cuda()) for t, x in zip(ts, xs))
    u0 = rng.random((9, 9, 9)).astype(np.float32)
    ud = D.f_empty(u0.shape, np.float32, "cuda")
    ud.copy_(torch.from_numpy(u0))
    itp = D.NDInterpolation(ud, dims)
    out = D.eval_grid(itp)
    assert out.is_cuda
    H.assert_close(out.cpu().numpy(), O.evaluate(["linear"] * 3, ts, u0, xs, grid=True), np.float32)
    # "repeated with new u": same caches, new data written into the borrowed device buffer
    u1 = rng.random((9, 9, 9)).astype(np.float32)
    ud.copy_(torch.from_numpy(u1))
    out2 = D.f_empty(out.shape, np.float32, "cuda")
    D.eval_grid_(out2, itp)
    H.assert_close(out2.cpu().numpy(), O.evaluate(["linear"] * 3, ts, u1, xs, grid=True), np.float32)
    itp.close()


# ---- cell-sorted plan: same numbers, original output positions ---------------------------------------


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [1, 3, 5, 9, 11])
def test_cell_sorted_plan_is_bit_identical(D, dtype, case):
    kinds, degrees, n_knots, out_dims, orders_list = CASES[case]
    rng = np.random.default_rng(400 + case)
    ts, u, _ = _case(rng, kinds, degrees, n_knots, out_dims, dtype)
    xs = H.zipped_queries(rng, ts, 20000, dtype)
    dims = H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[1] * len(kinds))
    itp = D.NDInterpolation(u, dims)
    for orders in orders_list:
        if any(o > 1 for o in orders):
            continue
        if "constant" in kinds and any(o > 0 for o in orders) and out_dims:
            a = np.full((xs[0].size,) + tuple(out_dims), 7.25, dtype=dtype, order="F")
            b = a.copy(order="F")
            D.eval_unstructured_(a, itp, derivative_orders=orders)
            D.eval_unstructured_(b, itp, derivative_orders=orders, flags=CELL_SORTED)
        else:
            a = D.eval_unstructured(itp, derivative_orders=orders)
            b = D.eval_unstructured(itp, derivative_orders=orders, flags=CELL_SORTED)
        assert np.array_equal(a, b, equal_nan=True)
    # a new u of the same shape re-uses the plan; new t_eval rebuilds it
    u2 = rng.random(u.shape).astype(dtype)
    itp.set_u(u2)
    a = D.eval_unstructured(itp)
    b = D.eval_unstructured(itp, flags=CELL_SORTED)
    assert np.array_equal(a, b, equal_nan=True)
    itp.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [1, 2, 3, 7, 9, 12])
def test_cell_sorted_plan_packed_cells_and_search_agree(D, dtype, case, monkeypatch):
    """The plan keeps its sort key — the per-dimension first nodes packed into one word — and the kernels decode
    the interval from it instead of searching; with DIND_NO_CELLS the plan sorts by the offset in u and the
    kernels search.  Same bits either way (value, derivative, fused gradient), equal to the unsorted result."""
    kinds, degrees, n_knots, out_dims, _ = CASES[case]
    rng = np.random.default_rng(900 + case)
    ts, u, _ = _case(rng, kinds, degrees, n_knots, out_dims, dtype)
    xs = H.zipped_queries(rng, ts, 30011, dtype)
    mk = lambda: D.NDInterpolation(u, H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[1] * len(kinds)))
    a, b, c = mk(), mk(), mk()
    orders = (1,) + (0,) * (len(kinds) - 1)
    ref = [D.eval_unstructured(a), D.eval_unstructured(a, derivative_orders=orders), D.eval_unstructured_gradient(a)]
    got = [D.eval_unstructured(b, flags=CELL_SORTED), D.eval_unstructured(b, derivative_orders=orders, flags=CELL_SORTED),
           D.eval_unstructured_gradient(b, flags=CELL_SORTED)]
    monkeypatch.setenv("DIND_NO_CELLS", "1")
    alt = [D.eval_unstructured(c, flags=CELL_SORTED), D.eval_unstructured(c, derivative_orders=orders, flags=CELL_SORTED),
           D.eval_unstructured_gradient(c, flags=CELL_SORTED)]
    monkeypatch.delenv("DIND_NO_CELLS")
    for r, g, l in zip(ref, got, alt):
        assert np.array_equal(r, g, equal_nan=True) and np.array_equal(r, l, equal_nan=True)
    assert np.array_equal(b.plan_permutation(), c.plan_permutation())   # same cell order from both keys
    for i in (a, b, c):
        i.close()


@pytest.mark.parametrize("tile", ["0", "2", "4"])
@pytest.mark.parametrize("shape", ["cubic3d", "linear5d"])
def test_tile_kernel_variants_are_bit_identical(D, tile, shape, monkeypatch):
    """DIND_C3_TILE selects the points per work item of the plan's tile kernels (0: one point per thread; 2 / 3 for
    the cubic kernel; 2 / 4 for the packed-FMA linear kernel).  Every variant must give the unsorted evaluation's bits,
    in original and in plan order, for values and a first derivative, and the fused gradient likewise."""
    rng = np.random.default_rng(1234)
    if shape == "cubic3d":
        kinds, degrees, n_knots, out_dims, dtype = ("bspline",) * 3, (3, 3, 3), (7, 6, 8), (3,), np.float64
    else:
        kinds, degrees, n_knots, out_dims, dtype = ("linear",) * 5, (0,) * 5, (4, 5, 4, 3, 5), (4,), np.float32
    ts, u, _ = _case(rng, kinds, degrees, n_knots, out_dims, dtype)
    xs = H.zipped_queries(rng, ts, 40000, dtype)      # ~100 points per cell: items of every length
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, np.finfo(dtype).max], dtype=dtype)
    xs = [np.concatenate([x, np.roll(special, d)]).astype(dtype) for d, x in enumerate(xs)]   # non-finite coordinates too
    n = len(kinds)
    mk = lambda: D.NDInterpolation(u, H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[1] * n))
    ref_itp = mk()
    monkeypatch.setenv("DIND_C3_TILE", tile)
    itp = mk()                                        # the knob is read once, in dind_create
    monkeypatch.delenv("DIND_C3_TILE")
    orders = (0,) * (n - 1) + (1,)
    for o in (None, orders):
        ref = D.eval_unstructured(ref_itp, derivative_orders=o)
        got = D.eval_unstructured(itp, derivative_orders=o, flags=CELL_SORTED)
        assert np.array_equal(ref, got, equal_nan=True), (o, itp.last_kernel)
        po = D.eval_unstructured(itp, derivative_orders=o, flags=CELL_SORTED | 32)
        assert np.array_equal(ref[itp.plan_permutation()], po, equal_nan=True), (o, itp.last_kernel)
        expect = ("unstructured_aos" if tile == "0" else ("unstructured_tile" if shape == "cubic3d" else "linear_tile"))
        assert itp.last_kernel.startswith(expect), itp.last_kernel
    assert np.array_equal(D.eval_unstructured_gradient(ref_itp), D.eval_unstructured_gradient(itp, flags=CELL_SORTED), equal_nan=True)
    # and against the oracle (values), NaN pattern included
    H.assert_close(D.eval_unstructured(itp, flags=CELL_SORTED), H.oracle_eval(kinds, ts, u, xs, degrees=degrees), dtype)
    itp.close()
    ref_itp.close()


@pytest.mark.parametrize("smem", ["0", "1"])
@pytest.mark.parametrize("n_points,n_comp", [(60000, 3), (9000, 3), (1500, 3), (20000, 4), (128, 3), (1, 3)])
def test_cubic_tile_kernel_shared_memory_staging(D, n_points, n_comp, smem, monkeypatch):
    """DIND_TILE_SMEM=1: the cubic tile kernel stages every distinct cell of a CTA's items in shared memory (cp.async) and
    contracts from there; more distinct cells than slots (sparse plans: 9000 / 1500 points over 1320 cells) go in rounds.
    Bit-identical to the unsorted evaluation and to the global-memory variant, values and derivatives, both point orders."""
    rng = np.random.default_rng(77 + n_points + n_comp)
    kinds, degrees, n_knots = ("bspline",) * 3, (3, 3, 3), (12, 11, 13)
    ts, u, _ = _case(rng, kinds, degrees, n_knots, (n_comp,), np.float64)
    xs = [rng.uniform(t[0] - 0.05, t[-1] + 0.05, n_points) for t in ts]
    mk = lambda: D.NDInterpolation(u, H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[1] * 3))
    ref_itp = mk()
    monkeypatch.setenv("DIND_TILE_SMEM", smem)
    itp = mk()
    monkeypatch.delenv("DIND_TILE_SMEM")
    for o in (None, (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0)):
        ref = D.eval_unstructured(ref_itp, derivative_orders=o)
        got = D.eval_unstructured(itp, derivative_orders=o, flags=CELL_SORTED)
        assert np.array_equal(ref, got, equal_nan=True), (o, itp.last_kernel)
        po = D.eval_unstructured(itp, derivative_orders=o, flags=CELL_SORTED | 32)
        assert np.array_equal(ref[itp.plan_permutation()], po, equal_nan=True), (o, itp.last_kernel)
    gref = D.eval_unstructured_gradient(ref_itp)
    assert np.array_equal(gref, D.eval_unstructured_gradient(itp, flags=CELL_SORTED), equal_nan=True), itp.last_kernel
    assert np.array_equal(gref[itp.plan_permutation()], D.eval_unstructured_gradient(itp, flags=CELL_SORTED | 32), equal_nan=True)
    H.assert_close(D.eval_unstructured(itp, flags=CELL_SORTED), H.oracle_eval(kinds, ts, u, xs, degrees=degrees), np.float64)
    itp.close()
    ref_itp.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_in,n_comp", [(2, 2), (3, 3), (3, 4), (4, 2), (4, 4), (5, 4), (5, 2), (6, 3)])
def test_cell_record_table_is_bit_identical(D, n_in, n_comp, dtype, monkeypatch):
    """DIND_REC=1 forces the cell-record copy of u (dind_fast_unstructured_rec.cu: every point's 2^N corners in one
    contiguous block; chosen by size for tables larger than L2, e.g. BASELINE config 5).  Same weights, same FMA order:
    the bits of the plain gather, for values, a derivative and with the cached idx_eval, incl. non-finite and
    out-of-span coordinates; and the oracle's values."""
    rng = np.random.default_rng(4321 + 10 * n_in + n_comp)
    # the kernel exists where a lane's share of a 128-byte line is whole nodes of one half-record
    has_rec = n_in >= 3 and (1 << (n_in - 1)) * (4 if n_comp == 3 else n_comp) * np.dtype(dtype).itemsize >= 128
    kinds, degrees = ("linear",) * n_in, (0,) * n_in
    n_knots = tuple([5, 4, 3, 6, 2, 3][:n_in])           # a 2-knot axis: one cell
    ts, u, _ = _case(rng, kinds, degrees, n_knots, (n_comp,), dtype)
    xs = H.zipped_queries(rng, ts, 20000, dtype)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, np.finfo(dtype).max], dtype=dtype)
    xs = [np.concatenate([x, np.roll(special, d)]).astype(dtype) for d, x in enumerate(xs)]
    mk = lambda: D.NDInterpolation(u, H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[1] * n_in))
    monkeypatch.setenv("DIND_REC", "0")
    ref_itp = mk()
    monkeypatch.setenv("DIND_REC", "1")
    itp = mk()                                        # the knob is read once, in dind_create
    monkeypatch.delenv("DIND_REC")
    for o in (None, (0,) * (n_in - 1) + (1,), (1,) + (0,) * (n_in - 1)):
        for flags in (0, 1):                          # 1 = DIND_FLAG_CACHED_IDX
            ref = D.eval_unstructured(ref_itp, derivative_orders=o, flags=flags)
            assert ref_itp.last_kernel.startswith("unstructured_aos"), ref_itp.last_kernel
            got = D.eval_unstructured(itp, derivative_orders=o, flags=flags)
            assert itp.last_kernel.startswith("unstructured_rec" if has_rec else "unstructured_aos"), itp.last_kernel
            assert np.array_equal(ref, got, equal_nan=True), (o, flags)
    # a new u rebuilds the table
    u2 = (u * 0.5 + 1).astype(dtype)
    itp.set_u(u2)
    ref_itp.set_u(u2)
    assert np.array_equal(D.eval_unstructured(ref_itp), D.eval_unstructured(itp), equal_nan=True)
    assert itp.last_kernel.startswith("unstructured_rec" if has_rec else "unstructured_aos")
    # the oracle's values (floatmax coordinates overflow to inf - inf in another order there: left to the bit comparison)
    ok = np.all([np.abs(np.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)) < 1e30 for x in xs], axis=0)
    H.assert_close(D.eval_unstructured(itp)[ok], H.oracle_eval(kinds, ts, u2, [x[ok] for x in xs], degrees=degrees), dtype)
    itp.close()
    ref_itp.close()


def test_nurbs_weights_with_constant_axis(D):
    """BASELINE config 4 names NURBS weights with mixed Constant(left=true)/BSpline dimensions: no
    reference semantics — defined as the tensor product of the per-dimension stencils
    with the weights multiplied in, checked against the oracle's mixed evaluator and by slice
    equivalence against the all-BSpline NURBS evaluation of the selected slice."""
    rng = np.random.default_rng(77)
    kinds, degrees = ("bspline", "bspline", "constant", "bspline"), (2, 2, 0, 2)
    ts, u, w = _case(rng, kinds, degrees, (6, 5, 5, 6), (), np.float64, nurbs=True)
    xg = [np.sort(H.query(rng, t, 9, np.float64, outside=False)) for t in ts]
    ref = H.oracle_eval(kinds, ts, u, xg, degrees=degrees, weights=w, grid=True)
    for flags in (0, GENERIC):
        got, kern = H.gpu_eval(D, kinds, ts, u, xg, degrees=degrees, weights=w, grid=True, flags=flags)
        H.assert_close(got, ref, np.float64)
    j = np.where(xg[2] >= ts[2][-1], len(ts[2]), O.get_idx("constant", ts[2], xg[2])) - 1
    for k in (0, len(xg[2]) // 2, len(xg[2]) - 1):
        sl = H.oracle_eval(("bspline",) * 3, [ts[0], ts[1], ts[3]], u[:, :, j[k], :], [xg[0], xg[1], xg[3]],
                           degrees=(2, 2, 2), weights=w[:, :, j[k], :], grid=True)
        H.assert_close(got[:, :, k, :], sl, np.float64)


# ---- fused value + gradient (beyond the reference: SURVEY.md §8f rank 1) ---------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("flags", [0, CELL_SORTED, GENERIC])
@pytest.mark.parametrize("case", [1, 2, 3, 7, 9, 12])
def test_fused_gradient_equals_separate_evaluations(D, dtype, flags, case):
    kinds, degrees, n_knots, out_dims, _ = CASES[case]
    rng = np.random.default_rng(500 + case)
    ts, u, _ = _case(rng, kinds, degrees, n_knots, out_dims, dtype)
    xs = H.zipped_queries(rng, ts, 4000, dtype)
    n = len(kinds)
    dims = H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[1] * n)
    itp = D.NDInterpolation(u, dims)
    g = D.eval_unstructured_gradient(itp, flags=flags)
    assert g.shape == (xs[0].size,) + tuple(out_dims) + (n + 1,)
    for s in range(n + 1):
        orders = tuple(1 if d + 1 == s else 0 for d in range(n))
        ref = H.oracle_eval(kinds, ts, u, xs, degrees=degrees, derivative_orders=orders)
        H.assert_close(g[..., s], ref, dtype)
    itp.close()


def test_fused_gradient_rejects_constant_and_nurbs(D):
    t = np.array([0.0, 1.0, 2.0, 3.0])
    itp = D.NDInterpolation(np.zeros(4), D.ConstantInterpolationDimension(t, t_eval=[0.5]))
    with pytest.raises(AssertionError):
        D.eval_unstructured_gradient(itp)
    itp.close()
    bs = D.BSplineInterpolationDimension(t, 2, t_eval=[0.5], max_derivative_order_eval=1)
    itp = D.NDInterpolation(np.zeros(5), bs, cache=D.NURBSWeights(np.ones(5)))
    with pytest.raises(AssertionError):
        D.eval_unstructured_gradient(itp)
    itp.close()


def test_plan_order_output_and_permutation(D):
    rng = np.random.default_rng(9)
    kinds, degrees = ("bspline",) * 2, (3, 2)
    ts, u, _ = _case(rng, kinds, degrees, (9, 8), (2,), np.float64)
    xs = H.zipped_queries(rng, ts, 5000, np.float64)
    itp = D.NDInterpolation(u, H.make_dims(D, kinds, ts, xs, degrees, None, max_deriv=[0, 0]))
    ref = D.eval_unstructured(itp)
    got = D.eval_unstructured(itp, flags=CELL_SORTED | 32)      # DIND_FLAG_PLAN_ORDER
    perm = itp.plan_permutation()
    assert sorted(perm.tolist()) == list(range(xs[0].size))
    assert np.array_equal(got, ref[perm])
    itp.close()


def test_pipelined_async_host_buffers(D):
    """DIND_FLAG_ASYNC with pinned host u / out: copies are queued (D2H on the copy stream), results are
    valid after synchronize() and equal the synchronous calls."""
    import torch
    rng = np.random.default_rng(12)
    ts = [H.knots(rng, 20, np.float32) for _ in range(3)]
    xs = [np.linspace(t[0], t[-1], 48).astype(np.float32) for t in ts]
    us = [torch.rand((20, 20, 20)).pin_memory() for _ in range(4)]
    u_np = [u.numpy().transpose(2, 1, 0) for u in us]
    outs = [torch.empty((48, 48, 48)).pin_memory() for _ in range(2)]
    out_np = [o.numpy().transpose(2, 1, 0) for o in outs]
    itp = D.NDInterpolation(u_np[0], tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs)))
    ref = []
    for k in range(4):
        itp.set_u(u_np[k])
        ref.append(np.array(D.eval_grid(itp)))
    got = []
    for k in range(4):
        itp.set_u(u_np[k], flags=1)
        D.eval_grid_(out_np[k % 2], itp, flags=1)
        if k % 2 == 1:          # both buffers in flight: drain and check
            itp.synchronize()
            got += [np.array(out_np[0]), np.array(out_np[1])]
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    itp.close()
