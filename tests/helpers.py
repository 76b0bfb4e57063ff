"""Shared helpers: build the same interpolation for the CUDA library (through the host mirror ->
C ABI) and for the CPU oracle, from one plain description."""
import numpy as np

from oracle import oracle as O

# north_star tolerances: Float64 1e-12 relative, Float32 1e-5 relative.  Applied norm-wise:
# |gpu - oracle| <= TOL * max|oracle| over the batch, because derivative results of smooth
# data cancel to ~0 point-wise.  The GPU contracts with FMA and uses
# reciprocal knot-span tables; the oracle rounds like Julia (no FMA, true divisions).
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def assert_close(got, ref, dtype, factor=1.0):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), "NaN pattern differs"
    scale = np.max(np.abs(ref[~nan_r])) if np.any(~nan_r) else 1.0
    scale = max(scale, np.finfo(dtype).tiny)
    err = np.max(np.abs(got[~nan_r] - ref[~nan_r])) if np.any(~nan_r) else 0.0
    assert err <= factor * TOL[np.dtype(dtype)] * scale, f"max err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


def make_dims(D, kinds, ts, t_evals, degrees=None, multiplicities=None, max_deriv=None):
    dims = []
    for i, k in enumerate(kinds):
        te = None if t_evals is None else t_evals[i]
        if k == "linear":
            dims.append(D.LinearInterpolationDimension(ts[i], t_eval=te))
        elif k == "constant":
            dims.append(D.ConstantInterpolationDimension(ts[i], t_eval=te))
        else:
            dims.append(D.BSplineInterpolationDimension(
                ts[i], degrees[i], t_eval=te,
                max_derivative_order_eval=0 if max_deriv is None else max_deriv[i],
                multiplicities=None if multiplicities is None else multiplicities[i]))
    return tuple(dims)


def gpu_eval(D, kinds, ts, u, t_evals, *, degrees=None, multiplicities=None, weights=None, derivative_orders=None,
             grid=False, flags=0, out=None):
    n = len(kinds)
    orders = derivative_orders or (0,) * n
    dims = make_dims(D, kinds, ts, t_evals, degrees, multiplicities, max_deriv=orders)
    itp = D.NDInterpolation(u, dims, cache=None if weights is None else D.NURBSWeights(weights))
    try:
        if out is not None:
            fn = D.eval_grid_ if grid else D.eval_unstructured_
            res = fn(out, itp, derivative_orders=derivative_orders, flags=flags)
        else:
            fn = D.eval_grid if grid else D.eval_unstructured
            res = fn(itp, derivative_orders=derivative_orders, flags=flags)
        return np.array(res), itp.last_kernel
    finally:
        itp.close()


def oracle_eval(kinds, ts, u, t_evals, **kw):
    kw.pop("flags", None)
    return O.evaluate(kinds, ts, u, t_evals, **kw)


def knots(rng, n, dtype=np.float64, uniform=False):
    """Non-uniform knots like docs/src/usage.md:11-12 (cumsum(0.5 .+ rand(n))), or uniform."""
    if uniform:
        return np.linspace(-1.0, 2.0, n).astype(dtype)
    return np.cumsum(0.5 + rng.random(n)).astype(dtype)


def query(rng, t, n, dtype=np.float64, with_knots=True, outside=True):
    """iid points in the domain + ~1 % outside + every knot value itself."""
    t = np.asarray(t, dtype=np.float64)
    pts = [rng.uniform(t[0], t[-1], n)]
    if outside:
        m = max(2, n // 100)
        span = t[-1] - t[0]
        pts.append(np.concatenate([rng.uniform(t[0] - 0.3 * span, t[0], m // 2 + 1), rng.uniform(t[-1], t[-1] + 0.3 * span, m // 2 + 1)]))
    if with_knots:
        pts.append(t)
    x = np.concatenate(pts)
    rng.shuffle(x)
    return x.astype(dtype)


def zipped_queries(rng, ts, n, dtype, outside=True):
    """Equal-length per-dimension query vectors (each including knots / outside points)."""
    xs = [query(rng, t, n, dtype, outside=outside) for t in ts]
    m = min(x.size for x in xs)
    return [x[:m] for x in xs]
