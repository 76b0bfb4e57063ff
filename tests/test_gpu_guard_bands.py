"""Out-of-bounds detection without compute-sanitizer (closed on this GPU pool): every device array the library
touches — u, weights, t_eval, out — sits in the middle of a larger allocation.  The guard bands around the inputs
are NaN, so a read outside an input poisons the result (caught by comparing with the evaluation of plain,
unguarded arrays bit for bit); the guard bands around `out` hold a canary that must survive.  Covers every kernel
family: pencil and two-stage grids, iid / plan-sorted / plan-order unstructured evaluation with the work-item tile
kernels (shared-memory staging on and off), the cell-record gather, the fused gradient, the generic fallback, and
stencils that touch the first and last nodes of u."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = 4096            # guard elements on either side
CANARY = -12345.678
CELL_SORTED, PLAN_ORDER, GENERIC = 16, 32, 8


@pytest.fixture(scope="module")
def D():
    import dind_b200
    assert dind_b200._lib.load().dind_device_count() > 0
    return dind_b200


def guarded(torch, a, fill):
    """column-major device copy of numpy array `a` inside a buffer with G guard elements of `fill` on both sides"""
    tdt = torch.float32 if a.dtype == np.float32 else torch.float64
    flat = torch.full((a.size + 2 * G,), fill, dtype=tdt, device="cuda")
    core = flat[G:G + a.size]
    core.copy_(torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))).cuda())
    nd = a.ndim
    view = core.view(a.shape[::-1]).permute(*reversed(range(nd))) if nd > 1 else core
    return flat, view


def plain(torch, D, a):
    t = D.f_empty(a.shape, a.dtype, "cuda")
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    return t


def knots(rng, n, dtype):
    return np.cumsum(0.5 + rng.random(n)).astype(dtype)


def points(rng, t, n, dtype):
    """in the span, outside it on both sides, and exactly on the first / last knots (stencils at the ends of u)"""
    span = float(t[-1] - t[0])
    x = np.concatenate([rng.uniform(t[0], t[-1], n - 40), rng.uniform(t[0] - 0.2 * span, t[0], 10),
                        rng.uniform(t[-1], t[-1] + 0.2 * span, 10), np.repeat(t[[0, -1]], 10)])
    rng.shuffle(x)
    return x.astype(dtype)


def make(D, kinds, degrees, ts, xs, u, w=None):
    dims = []
    for k, p, t, x in zip(kinds, degrees, ts, xs):
        if k == "linear":
            dims.append(D.LinearInterpolationDimension(t, t_eval=x))
        elif k == "constant":
            dims.append(D.ConstantInterpolationDimension(t, t_eval=x))
        else:
            dims.append(D.BSplineInterpolationDimension(t, p, t_eval=x, max_derivative_order_eval=1))
    return D.NDInterpolation(u, tuple(dims), cache=None if w is None else D.NURBSWeights(w))


UNSTRUCTURED = [
    # kinds, degrees, knots per dim, out dims, dtype, n points
    (("bspline",) * 3, (3, 3, 3), (9, 8, 10), (3,), np.float64, 60_000),      # config-3 shape: tile kernels (PT = 3 / gradient PT = 2)
    (("linear",) * 5, (0,) * 5, (5, 4, 5, 4, 6), (4,), np.float32, 60_000),   # config-5 shape: packed-FMA linear tile kernel
    (("linear",) * 2, (0, 0), (9, 7), (2,), np.float64, 20_000),              # config-1 shape
    (("bspline",) * 2, (2, 3), (7, 6), (), np.float32, 20_000),
    (("constant", "linear", "bspline"), (0, 0, 2), (5, 6, 7), (2,), np.float64, 5_000),   # mixed kinds: generic kernel
]


@pytest.mark.parametrize("variant", ["default", "cell-record table, global-memory tile kernels"])
@pytest.mark.parametrize("case", range(len(UNSTRUCTURED)))
def test_unstructured_kernels_stay_inside_their_arrays(D, case, variant, monkeypatch):
    import torch
    kinds, degrees, n_knots, out_dims, dtype, n = UNSTRUCTURED[case]
    rng = np.random.default_rng(1000 + case)
    ts = [knots(rng, k, dtype) for k in n_knots]
    n_u = [t.size + p - 1 if k == "bspline" else t.size for k, p, t in zip(kinds, degrees, ts)]
    u = rng.random(tuple(n_u) + tuple(out_dims)).astype(dtype)
    xs = [points(rng, t, n, dtype) for t in ts]
    ref_itp = make(D, kinds, degrees, ts, [plain(torch, D, x) for x in xs], plain(torch, D, u))
    keep = [guarded(torch, x, float("nan")) for x in xs] + [guarded(torch, u, float("nan"))]
    if variant != "default":
        # the guarded handle gathers multilinear tables from the cell-record copy of u (dind_fast_unstructured_rec.cu)
        # and runs the cubic tile kernels without their shared-memory staging; the reference handle keeps the defaults
        # (small table: plain gather; dense plan: shared-memory staging) — same bits expected
        monkeypatch.setenv("DIND_REC", "1")
        monkeypatch.setenv("DIND_TILE_SMEM", "0")
    itp = make(D, kinds, degrees, ts, [k[1] for k in keep[:-1]], keep[-1][1])
    monkeypatch.delenv("DIND_REC", raising=False)
    monkeypatch.delenv("DIND_TILE_SMEM", raising=False)
    n_in = len(kinds)
    orders_list = [None, tuple(1 if d == n_in - 1 else 0 for d in range(n_in))]
    if "constant" in kinds:
        orders_list = [None]
    shape = (n,) + tuple(out_dims)
    for orders in orders_list:
        for flags in (0, CELL_SORTED, CELL_SORTED | PLAN_ORDER, GENERIC):
            ref = D.eval_unstructured(ref_itp, derivative_orders=orders, flags=flags)
            flat, out = guarded(torch, np.zeros(shape, dtype=dtype), CANARY)
            D.eval_unstructured_(out, itp, derivative_orders=orders, flags=flags)
            assert torch.equal(out, ref), (orders, flags, itp.last_kernel)
            assert bool((flat[:G] == CANARY).all()) and bool((flat[-G:] == CANARY).all()), (orders, flags, itp.last_kernel)
    if "constant" not in kinds:
        gshape = shape + (n_in + 1,)
        for flags in (0, CELL_SORTED, CELL_SORTED | PLAN_ORDER):
            ref = D.eval_unstructured_gradient(ref_itp, flags=flags)
            flat, out = guarded(torch, np.zeros(gshape, dtype=dtype), CANARY)
            D.eval_unstructured_gradient(itp, out, flags=flags)
            assert torch.equal(out, ref), (flags, itp.last_kernel)
            assert bool((flat[:G] == CANARY).all()) and bool((flat[-G:] == CANARY).all()), (flags, itp.last_kernel)
    itp.close()
    ref_itp.close()


GRIDS = [
    (("linear",) * 3, (0,) * 3, (20, 18, 22), (), np.float32, (64, 40, 48), False),                 # config-2 shape: pencil kernel
    (("bspline",) * 4, (2,) * 4, (9, 8, 9, 8), (), np.float64, (64, 32, 20, 24), True),               # config-4 shape: two-stage NURBS
    (("bspline", "bspline", "constant", "bspline"), (2, 2, 0, 2), (9, 8, 7, 8), (), np.float64, (64, 32, 9, 24), True),
    (("bspline",) * 2, (3, 3), (12, 11), (2,), np.float64, (70, 33), False),
    (("constant",) * 3, (0,) * 3, (8, 9, 7), (), np.float32, (40, 30, 20), False),
]


@pytest.mark.parametrize("persist", [False, True, "scalar batch loads"])
@pytest.mark.parametrize("case", range(len(GRIDS)))
def test_grid_kernels_stay_inside_their_arrays(D, case, persist, monkeypatch):
    import torch
    kinds, degrees, n_knots, out_dims, dtype, n_g, nurbs = GRIDS[case]
    if persist and case not in ((1,) if persist is True else (1, 2)):
        pytest.skip("variants of the two-stage 4-D kernels only")
    if persist is True:
        monkeypatch.setenv("DIND_PERSIST", "1")
    elif persist:
        monkeypatch.setenv("DIND_NO_BATCH_VEC", "1")      # stage 2 with one scalar load per batch column (the odd-B path)
    rng = np.random.default_rng(2000 + case)
    ts = [knots(rng, k, dtype) for k in n_knots]
    n_u = [t.size + p - 1 if k == "bspline" else t.size for k, p, t in zip(kinds, degrees, ts)]
    u = rng.random(tuple(n_u) + tuple(out_dims)).astype(dtype)
    w = (0.5 + rng.random(tuple(n_u))).astype(dtype) if nurbs else None
    xs = []
    for t, g in zip(ts, n_g):
        x = np.sort(rng.uniform(t[0], t[-1], g)).astype(dtype)
        x[0], x[-1] = t[0], t[-1]
        xs.append(x)
    ref_itp = make(D, kinds, degrees, ts, [plain(torch, D, x) for x in xs], plain(torch, D, u), None if w is None else plain(torch, D, w))
    keep = [guarded(torch, x, float("nan")) for x in xs] + [guarded(torch, u, float("nan"))]
    wk = guarded(torch, w, float("nan")) if nurbs else None
    itp = make(D, kinds, degrees, ts, [k[1] for k in keep[:-1]], keep[-1][1], None if wk is None else wk[1])
    shape = tuple(n_g) + tuple(out_dims)
    for flags in (0, GENERIC):
        ref = D.eval_grid(ref_itp, flags=flags)
        flat, out = guarded(torch, np.zeros(shape, dtype=dtype), CANARY)
        D.eval_grid_(out, itp, flags=flags)
        assert torch.equal(out, ref), (flags, itp.last_kernel)
        assert bool((flat[:G] == CANARY).all()) and bool((flat[-G:] == CANARY).all()), (flags, itp.last_kernel)
    itp.close()
    ref_itp.close()
