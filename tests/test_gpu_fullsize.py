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


# ---- config 3: 3-D cubic BSpline, u (128,128,128,3) Float64, 10^8 zipped points + first derivatives --------
def test_c3_full_size_1e8_points(D, torch):
    """BASELINE configs[2] at its named size: 10^8 points, value and the three first partials.  Points come from the
    counter-based generator (tests/helpers.py) so the oracle regenerates any subsample; 1 % of the span on either
    side extrapolates."""
    from tests import helpers as H
    rng = np.random.default_rng(3)
    ts = [knots(rng, 126, np.float64) for _ in range(3)]
    n = 100_000_000
    xs_d = H.counter_points_torch(0, n, ts, 31, np.float64, "cuda", overshoot=0.01)
    dims = tuple(D.BSplineInterpolationDimension(t, 3, t_eval=x, max_derivative_order_eval=1) for t, x in zip(ts, xs_d))
    shape = (128, 128, 128, 3)
    gen = torch.Generator(device="cuda").manual_seed(7)
    u = D.f_empty(shape, np.float64, "cuda")
    u.copy_(torch.rand(shape[::-1], generator=gen, device="cuda", dtype=torch.float64).permute(3, 2, 1, 0))
    itp = D.NDInterpolation(u, dims)
    orders_all = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1))
    sel_h = rng.integers(0, n, 100_000)
    sel = torch.from_numpy(sel_h).cuda()
    xs_h = H.counter_points_np(sel_h, ts, 31, np.float64, overshoot=0.01)
    for d in range(3):   # the device generator and the numpy one agree bit for bit
        assert np.array_equal(xs_d[d][sel].cpu().numpy(), xs_h[d])
    u_h = u.cpu().numpy()
    out = D.f_empty((n, 3), np.float64, "cuda")
    out_s = D.f_empty((n, 3), np.float64, "cuda")
    worst = 0.0
    for o in orders_all:
        D.eval_unstructured_(out, itp, derivative_orders=o)
        assert itp.last_kernel.startswith("unstructured")
        ref = O.evaluate(["bspline"] * 3, ts, u_h, xs_h, degrees=[3] * 3, derivative_orders=o)
        got = out[sel].cpu().numpy()
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        worst = max(worst, err)
        assert err <= 1e-12, (o, err)
        # the cell-sorted plan gives bit-identical results at the original positions — all 10^8 of them
        D.eval_unstructured_(out_s, itp, derivative_orders=o, flags=CELL_SORTED)
        assert torch.equal(out_s, out), o
    H.record_observed(worst / 1e-12, worst)
    # interval indices of the subsample, bit-exact (get_idx, src/interpolation_utils.jl:104-134)
    for d in range(3):
        sub = D.BSplineInterpolationDimension(ts[d], 3, t_eval=xs_h[d])
        assert np.array_equal(sub.idx_eval, O.get_idx("bspline", ts[d], xs_h[d], degree=3))
    # fused value + gradient == the four separate evaluations (same summation order per slot is not promised: 1e-12)
    del out_s
    g = D.eval_unstructured_gradient(itp, flags=CELL_SORTED)
    for s_, o in enumerate(orders_all):
        ref = O.evaluate(["bspline"] * 3, ts, u_h, xs_h, degrees=[3] * 3, derivative_orders=o)
        got = g[..., s_][sel].cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), o
    del g
    # partition of unity: a constant control net gives the constant, and zero gradient
    u.fill_(2.0)
    itp.set_u(u)
    D.eval_unstructured_(out, itp)
    assert float((out - 2.0).abs().max()) <= 1e-12
    D.eval_unstructured_(out, itp, derivative_orders=(0, 0, 1))
    assert float(out.abs().max()) <= 1e-11
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


def test_c4_persistent_two_stage_kernel_matches_two_launches(D, torch, monkeypatch):
    """DIND_PERSIST=1 runs both separable stages of the 4-D grid evaluation in one persistent launch (stage-2 CTAs
    produce the stage-1 planes they wait for; dind_fast_grid_split.cu).  Opt-in because it measured slower than the
    two launches; same numbers up to the summation order of the window, any t_eval order along the last axis."""
    rng = np.random.default_rng(44)
    ts = [knots(rng, 31, np.float64) for _ in range(4)]
    shape = (32,) * 4
    gen = torch.Generator(device="cuda").manual_seed(19)
    u = D.f_empty(shape, np.float64, "cuda"); u.copy_(torch.rand(shape, generator=gen, device="cuda", dtype=torch.float64))
    w = D.f_empty(shape, np.float64, "cuda"); w.copy_(torch.rand(shape, generator=gen, device="cuda", dtype=torch.float64) + 0.5)
    for sort_last in (True, False):
        xs = [np.sort(rng.uniform(t[0], t[-1], n)) for t, n in zip(ts, (64, 64, 40, 50))]
        if not sort_last:
            rng.shuffle(xs[3])
        for cache in (None, D.NURBSWeights(w)):
            outs = []
            for persist in (False, True):
                if persist:
                    monkeypatch.setenv("DIND_PERSIST", "1")
                else:
                    monkeypatch.delenv("DIND_PERSIST", raising=False)
                dims = tuple(D.BSplineInterpolationDimension(t, 2, t_eval=torch.from_numpy(x).cuda()) for t, x in zip(ts, xs))
                itp = D.NDInterpolation(u, dims, cache=cache)
                outs.append(D.eval_grid(itp))
                assert itp.last_kernel.startswith("grid_split_fused" if persist else "grid_split<"), itp.last_kernel
                itp.close()
            # the last-axis walk sums its window in slot order, and the slot <-> plane rotation depends on where a
            # chunk starts (8-output chunks here, 32 in the two-launch path): equal up to the summation order
            assert float((outs[0] - outs[1]).abs().max()) <= 1e-14 * float(outs[0].abs().max())
    monkeypatch.delenv("DIND_PERSIST", raising=False)


# ---- config 5: 5-D linear lookup table u (32^5, 4) Float32, point-sharded ------------------------------------
def _c5_setup(D, torch, seed=5):
    rng = np.random.default_rng(seed)
    ts = [knots(rng, 32, np.float32) for _ in range(5)]
    shape = (32,) * 5 + (4,)
    gen = torch.Generator(device="cuda").manual_seed(11)
    u = D.f_empty(shape, np.float32, "cuda")
    u.copy_(torch.rand(shape[::-1], generator=gen, device="cuda").permute(*reversed(range(6))))   # random u: every corner matters
    return rng, ts, u


def _c5_check_subsample(D, torch, H, ts, u_h, out, lo, sel_h, orders, seed):
    """oracle on the global points sel_h (lo <= sel_h < lo + len(out)): values at 1e-5, regenerated coordinates."""
    xs_h = H.counter_points_np(sel_h, ts, seed, np.float32, overshoot=0.01)
    ref = O.evaluate(["linear"] * 5, ts, u_h, xs_h, derivative_orders=orders)
    got = out[torch.from_numpy(sel_h - lo).cuda()].cpu().numpy()
    err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    assert err <= 1e-5, (orders, err)
    return err, xs_h


def test_c5_random_u_against_oracle_and_sharding(D, torch):
    """Random u (an affine u is reproduced from ANY cell, so it cannot see a wrong interval or corner order —
    src/interpolation_utils.jl:116-134, src/interpolation_methods.jl:22-36): value and one first derivative against
    the oracle on a 10^5-point subsample, interval indices of the subsample bit-exact, plan == no plan bit for bit,
    index-range shards tile the single-handle result."""
    from tests import helpers as H
    rng, ts, u = _c5_setup(D, torch)
    n, seed = 50_000_000, 41
    xs_d = H.counter_points_torch(0, n, ts, seed, np.float32, "cuda", overshoot=0.01)
    dims = tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_d))
    itp = D.NDInterpolation(u, dims)
    u_h = u.cpu().numpy()
    sel_h = rng.integers(0, n, 100_000)
    out = D.eval_unstructured(itp)
    e0, xs_h = _c5_check_subsample(D, torch, H, ts, u_h, out, 0, sel_h, (0,) * 5, seed)
    assert torch.equal(D.eval_unstructured(itp, flags=CELL_SORTED), out)
    d3 = D.eval_unstructured(itp, derivative_orders=(0, 0, 1, 0, 0))
    e1, _ = _c5_check_subsample(D, torch, H, ts, u_h, d3, 0, sel_h, (0, 0, 1, 0, 0), seed)
    assert torch.equal(D.eval_unstructured(itp, derivative_orders=(0, 0, 1, 0, 0), flags=CELL_SORTED), d3)
    H.record_observed(max(e0, e1) / 1e-5, max(e0, e1))
    for d in range(5):
        assert np.array_equal(xs_d[d][torch.from_numpy(sel_h).cuda()].cpu().numpy(), xs_h[d])
        sub = D.LinearInterpolationDimension(ts[d], t_eval=xs_h[d])
        assert np.array_equal(sub.idx_eval, O.get_idx("linear", ts[d], xs_h[d], dtype=np.float32))
    itp.close()
    # index-range sharding: rank r of 4 evaluates its slice with its own handle; the pieces tile the result
    for r in range(4):
        lo, hi = D.shard_range(n, r, 4)
        it = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x[lo:hi]) for t, x in zip(ts, xs_d)))
        assert torch.equal(D.eval_unstructured(it), out[lo:hi])
        it.close()


def test_c5_full_size_4e9_points_in_shards_on_one_gpu(D, torch):
    """BASELINE configs[4] at its named size, 4*10^9 points, on ONE GPU: the eight index-range shards an 8-GPU job
    would own (5*10^8 points each, generated on the device from the global index) are evaluated one after the
    other by one handle; every shard is checked against the oracle on a subsample and through the multilinear
    partition of unity (the four components of a constant table)."""
    from tests import helpers as H
    rng, ts, u = _c5_setup(D, torch)
    u_h = u.cpu().numpy()
    n_total, world, seed = 4_000_000_000, 8, 43
    itp = None
    worst = 0.0
    out = None
    for r in range(world):
        lo, hi = D.shard_range(n_total, r, world)
        xs_d = H.counter_points_torch(lo, hi, ts, seed, np.float32, "cuda", overshoot=0.01)
        if itp is None:
            itp = D.NDInterpolation(u, tuple(D.LinearInterpolationDimension(t, t_eval=x) for t, x in zip(ts, xs_d)))
            out = D.f_empty((hi - lo, 4), np.float32, "cuda")
        else:
            itp.set_t_eval(xs_d)
        D.eval_unstructured_(out, itp)
        sel_h = lo + rng.integers(0, hi - lo, 20_000)
        err, _ = _c5_check_subsample(D, torch, H, ts, u_h, out, lo, sel_h, (0,) * 5, seed)
        worst = max(worst, err)
        assert bool(torch.isfinite(out).all())
        del xs_d
    H.record_observed(worst / 1e-5, worst)
    itp.close()
