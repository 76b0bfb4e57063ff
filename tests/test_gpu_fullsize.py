"""Parity at BASELINE.json's full sizes (run on the B200 box: pytest -m gpu).

The oracle cannot evaluate 10^8 points in seconds, so at full size the CUDA path is checked through
size-independent properties of the interpolants — reproduction of constants and of (multi)linear
functions, linearity in u, slab equivalence, plan equivalence — plus an oracle comparison on a
random subsample.  All data stays on the device (torch is only the container)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CELL_SORTED = 16


@pytest.fixture(scope="module")
def D():
    import dind_b200
    assert dind_b200._lib.load().dind_device_count() > 0
    return dind_b200


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def knots(rng, n, dtype):
    return np.cumsum(0.5 + rng.random(n)).astype(dtype)


def dev_f(D, torch, a):
    t = D.f_empty(a.shape, a.dtype, "cuda")
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    return t


# ---- config 2: 3-D trilinear, u 256^3 Float32 -> 512^3 ------------------------------------------------
def test_c2_full_size_properties(D, torch):
    rng = np.random.default_rng(2)
    n_u, n_g = 256, 512
    ts = [knots(rng, n_u, np.float32) for _ in range(3)]
    xs = [np.linspace(t[0], t[-1], n_g).astype(np.float32) for t in ts]
    xs_d = [torch.from_numpy(x).cuda() for x in xs]
    dims = tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_d))
    gen = torch.Generator(device="cuda").manual_seed(5)
    u1 = D.f_empty((n_u,) * 3, np.float32, "cuda"); u1.copy_(torch.rand((n_u,) * 3, generator=gen, device="cuda"))
    u2 = D.f_empty((n_u,) * 3, np.float32, "cuda"); u2.copy_(torch.rand((n_u,) * 3, generator=gen, device="cuda"))
    itp = D.NDInterpolation(u1, dims)
    o1 = D.eval_grid(itp)
    assert itp.last_kernel.startswith(("grid_pencil", "grid_split"))
    assert tuple(o1.shape) == (n_g,) * 3
    # (a) oracle on a random subsample of grid points (as zipped points)
    idx = rng.integers(0, n_g, size=(3, 200_000))
    sub = o1[torch.from_numpy(idx[0]).cuda(), torch.from_numpy(idx[1]).cuda(), torch.from_numpy(idx[2]).cuda()].cpu().numpy()
    ref = O.evaluate(["linear"] * 3, ts, u1.cpu().numpy(), [xs[d][idx[d]] for d in range(3)])
    assert np.max(np.abs(sub - ref)) <= 1e-5 * np.max(np.abs(ref))
    # (b) linearity in u
    o2 = D.eval_grid(itp.set_u(u2))
    u3 = D.f_empty((n_u,) * 3, np.float32, "cuda"); u3.copy_(0.75 * u1 - 1.5 * u2)
    o3 = D.eval_grid(itp.set_u(u3))
    assert float((o3 - (0.75 * o1 - 1.5 * o2)).abs().max()) <= 1e-5 * 2.25
    # (c) reproduction of a constant and of a trilinear function of (t1, t2, t3)
    u3.fill_(2.0)
    assert float((D.eval_grid(itp.set_u(u3)) - 2.0).abs().max()) <= 1e-5 * 2.0
    t_d = [torch.from_numpy(t).cuda() for t in ts]
    f = lambda a, b, c: 3.0 + 0.023 * a - 0.047 * b + 0.011 * c
    u3.copy_(f(t_d[0][:, None, None], t_d[1][None, :, None], t_d[2][None, None, :]))
    exp = f(xs_d[0][:, None, None], xs_d[1][None, :, None], xs_d[2][None, None, :])
    got = D.eval_grid(itp.set_u(u3))
    assert float((got - exp).abs().max()) <= 2e-5 * float(exp.abs().max())
    # (d) derivative along axis 2 of that function is the constant -0.047 (same cached tables, other order)
    # (e) slab equivalence: two half slabs along the last axis == the full grid, bit for bit
    itp.set_u(u1)
    halves = []
    for lo, hi in ((0, n_g // 2), (n_g // 2, n_g)):
        d2 = tuple(D.LinearInterpolationDimension(t, t_eval=(x if k < 2 else x[lo:hi])) for k, (t, x) in enumerate(zip(ts, xs_d)))
        it2 = D.NDInterpolation(u1, d2)
        halves.append(D.eval_grid(it2))
        it2.close()
    assert torch.equal(torch.cat([h.permute(2, 1, 0) for h in halves], dim=0).permute(2, 1, 0), o1)
    itp.close()


def test_c2_first_derivative_of_trilinear_function(D, torch):
    rng = np.random.default_rng(3)
    ts = [knots(rng, 256, np.float32) for _ in range(3)]
    xs_d = [torch.from_numpy(np.linspace(t[0], t[-1], 512).astype(np.float32)).cuda() for t in ts]
    t_d = [torch.from_numpy(t).cuda() for t in ts]
    u = D.f_empty((256,) * 3, np.float32, "cuda")
    u.copy_(3.0 + 0.5 * t_d[0][:, None, None] - 0.25 * t_d[1][None, :, None] + 0.125 * t_d[2][None, None, :])
    itp = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_d)))
    for orders, slope in (((1, 0, 0), 0.5), ((0, 1, 0), -0.25), ((0, 0, 1), 0.125)):
        d = D.eval_grid(itp, derivative_orders=orders)
        assert float((d - slope).abs().max()) <= 2e-3 * abs(slope)   # F32 differences of O(100) values over O(1) spans
    assert float(D.eval_grid(itp, derivative_orders=(2, 0, 0)).abs().max()) == 0.0
    itp.close()


# ---- config 1: 2-D linear, u (64,64,2) Float64, 10^6 zipped points: full oracle comparison ------------
def test_c1_full_size_against_oracle(D, torch):
    rng = np.random.default_rng(1)
    ts = [knots(rng, 64, np.float64) for _ in range(2)]
    u = rng.random((64, 64, 2))
    xs = [rng.uniform(t[0] - 0.5, t[-1] + 0.5, 10 ** 6) for t in ts]
    itp = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs)))
    for orders in ((0, 0), (1, 0), (0, 1), (1, 1)):
        got = D.eval_unstructured(itp, derivative_orders=orders)
        ref = O.evaluate(["linear"] * 2, ts, u, xs, derivative_orders=orders)
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
        assert np.array_equal(got, D.eval_unstructured(itp, derivative_orders=orders, flags=CELL_SORTED))
    for d in range(2):
        assert np.array_equal(itp.interp_dims[d].idx_eval, O.get_idx("linear", ts[d], xs[d]))
    itp.close()


# ---- config 3: 3-D cubic BSpline, u (128,128,128,3) Float64 ----------------------------------------------
def test_c3_properties_and_subsample(D, torch):
    rng = np.random.default_rng(3)
    ts = [knots(rng, 126, np.float64) for _ in range(3)]
    n = 20_000_000
    gen = torch.Generator(device="cuda").manual_seed(7)
    xs_d = [torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) * float(t[-1] - t[0]) + float(t[0]) for t in ts]
    dims = tuple(D.BSplineInterpolationDimension(t, 3, t_eval=x, max_derivative_order_eval=1) for t, x in zip(ts, xs_d))
    shape = (128, 128, 128, 3)
    u = D.f_empty(shape, np.float64, "cuda")
    u.copy_(torch.rand(shape[::-1], generator=gen, device="cuda", dtype=torch.float64).permute(3, 2, 1, 0))
    itp = D.NDInterpolation(u, dims)
    orders_all = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1))
    outs = {o: D.eval_unstructured(itp, derivative_orders=o) for o in orders_all}
    assert itp.last_kernel.startswith("unstructured")
    # oracle on a subsample
    sel = torch.from_numpy(rng.integers(0, n, 100_000)).cuda()
    xs_h = [x[sel].cpu().numpy() for x in xs_d]
    u_h = u.cpu().numpy()
    for o in orders_all:
        ref = O.evaluate(["bspline"] * 3, ts, u_h, xs_h, degrees=[3] * 3, derivative_orders=o)
        got = outs[o][sel].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    # the cell-sorted plan gives bit-identical results at the original positions
    for o in ((0, 0, 0), (0, 1, 0)):
        assert torch.equal(D.eval_unstructured(itp, derivative_orders=o, flags=CELL_SORTED), outs[o])
    # partition of unity: a constant control net gives the constant, and zero gradient
    u.fill_(2.0)
    itp.set_u(u)
    assert float((D.eval_unstructured(itp) - 2.0).abs().max()) <= 1e-12
    assert float(D.eval_unstructured(itp, derivative_orders=(0, 0, 1)).abs().max()) <= 1e-11
    itp.close()


# ---- config 4: 4-D degree-2 NURBS, u 64^4 (+ weights) Float64 -> 128^4 ---------------------------------------
def test_c4_nurbs_grid_properties(D, torch):
    rng = np.random.default_rng(4)
    ts = [knots(rng, 63, np.float64) for _ in range(4)]
    xs = [np.linspace(t[0], t[-1], 128) for t in ts]
    gen = torch.Generator(device="cuda").manual_seed(9)
    shape = (64,) * 4
    u = D.f_empty(shape, np.float64, "cuda"); u.copy_(torch.rand(shape, generator=gen, device="cuda", dtype=torch.float64))
    w = D.f_empty(shape, np.float64, "cuda"); w.copy_(torch.rand(shape, generator=gen, device="cuda", dtype=torch.float64) + 0.5)
    dims = tuple(D.BSplineInterpolationDimension(t, 2, t_eval=torch.from_numpy(x).cuda()) for t, x in zip(ts, xs))
    itp = D.NDInterpolation(u, dims, cache=D.NURBSWeights(w))
    out = D.eval_grid(itp)
    assert tuple(out.shape) == (128,) * 4 and itp.last_kernel.startswith(("grid_pencil", "grid_split"))
    idx = rng.integers(0, 128, size=(4, 20_000))
    sub = out[tuple(torch.from_numpy(i).cuda() for i in idx)].cpu().numpy()
    ref = O.evaluate(["bspline"] * 4, ts, u.cpu().numpy(), [xs[d][idx[d]] for d in range(4)], degrees=[2] * 4, weights=w.cpu().numpy())
    assert np.max(np.abs(sub - ref)) <= 1e-12 * np.max(np.abs(ref))
    # a rational B-spline reproduces constants whatever the weights
    u.fill_(2.0)
    itp.set_u(u)
    assert float((D.eval_grid(itp) - 2.0).abs().max()) <= 1e-12
    itp.close()


# ---- config 5: 5-D linear lookup table u (32^5, 4) Float32, point-sharded ------------------------------------
def test_c5_multilinear_reproduction_and_sharding(D, torch):
    rng = np.random.default_rng(5)
    ts = [knots(rng, 32, np.float32) for _ in range(5)]
    n = 50_000_000
    gen = torch.Generator(device="cuda").manual_seed(11)
    xs_d = [torch.rand(n, generator=gen, device="cuda") * float(t[-1] - t[0]) + float(t[0]) for t in ts]
    t_d = [torch.from_numpy(t).cuda() for t in ts]
    coef = [0.3, -0.2, 0.15, 0.4, -0.1]
    shape = (32,) * 5 + (4,)
    u = D.f_empty(shape, np.float32, "cuda")
    base = 1.0
    for d in range(5):
        view = [1] * 5
        view[d] = 32
        base = base + coef[d] * t_d[d].reshape(view)
    for c in range(4):
        u[..., c] = (c + 1) * base     # component c is (c+1) times a function that is linear in every axis
    dims = tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_d))
    itp = D.NDInterpolation(u, dims)
    out = D.eval_unstructured(itp)
    exp = 1.0 + sum(cf * x for cf, x in zip(coef, xs_d))
    for c in range(4):
        assert float((out[:, c] - (c + 1) * exp).abs().max()) <= 2e-5 * (c + 1) * float(exp.abs().max())
    assert torch.equal(D.eval_unstructured(itp, flags=CELL_SORTED), out)
    itp.close()
    # index-range sharding: rank r of 4 evaluates its slice with its own handle; the pieces tile the result
    for r in range(4):
        lo, hi = D.shard_range(n, r, 4)
        it = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x[lo:hi]) for t, x in zip(ts, xs_d)))
        assert torch.equal(D.eval_unstructured(it), out[lo:hi])
        it.close()
