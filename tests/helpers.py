"""Shared helpers: build the same interpolation for the CUDA library (through the host mirror ->
C ABI) and for the CPU oracle, from one plain description."""
import numpy as np

from oracle import oracle as O

# north_star tolerances: Float64 1e-12 relative, Float32 1e-5 relative.  Applied norm-wise:
# |gpu - oracle| <= TOL * max|oracle| over the batch, because derivative results of smooth
# data cancel to ~0 point-wise.  The GPU contracts with FMA and uses
# reciprocal knot-span tables; the oracle rounds like Julia (no FMA, true divisions).
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


# observed error / tolerance of every assert_close call, keyed by the running test: tests/conftest.py writes
# the per-test maximum to gpurun_out/parity_observed.json (and profiles/ keeps a copy of the last GPU run)
OBSERVED = {}


def record_observed(ratio, rel):
    import os
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    prev = OBSERVED.get(test, (0.0, 0.0))
    OBSERVED[test] = (max(prev[0], float(ratio)), max(prev[1], float(rel)))


def assert_close(got, ref, dtype, factor=1.0):
    """|got - ref| <= factor * TOL * max|ref|.  `factor` stays 1 (the north-star tolerance) unless the call site
    gives a one-line reason; DIND_TOL_SOFT=1 records the observed ratio without failing (survey runs)."""
    import os
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert np.array_equal(nan_g, nan_r), "NaN pattern differs"
    scale = np.max(np.abs(ref[~nan_r])) if np.any(~nan_r) else 1.0
    scale = max(scale, np.finfo(dtype).tiny)
    err = np.max(np.abs(got[~nan_r] - ref[~nan_r])) if np.any(~nan_r) else 0.0
    record_observed(err / (TOL[np.dtype(dtype)] * scale), err / scale)
    if os.environ.get("DIND_TOL_SOFT"):
        return
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


# ---- counter-based RNG keyed by (seed, dim, GLOBAL point index) ------------------------------------------------
# anything >= 1e8 points is generated on the device, per shard, from the global index, so that any
# rank (or a single GPU, or the CPU oracle on a subsample) regenerates exactly the same point.  A 32-bit integer
# hash evaluated in int64 with the value masked to 32 bits after every step: numpy and torch give identical bits.
def counter_u01(idx, dim, seed):
    """idx: int64 numpy array or torch tensor of global indices -> float64 uniforms in [0, 1)."""
    M = 0xFFFFFFFF
    x = (idx & M) ^ (((idx >> 32) * 0x9E3779B1) & M)
    x = (x + ((dim * 0x85EBCA6B + seed * 0xC2B2AE35 + 0x27D4EB2F) & M)) & M
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & M
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & M
    x = x ^ (x >> 16)
    if hasattr(x, "numpy") or type(x).__module__.startswith("torch"):
        import torch
        return x.to(torch.float64) * (1.0 / 4294967296.0)
    return x.astype(np.float64) * (1.0 / 4294967296.0)


def counter_points_np(idx, ts, seed, dtype, overshoot=0.0):
    """Per-dimension coordinates of the global points `idx` (numpy int64): uniform over the knot span, stretched by
    `overshoot` of the span on both sides (so a fraction of the points extrapolates)."""
    idx = np.asarray(idx, dtype=np.int64)
    out = []
    for d, t in enumerate(ts):
        t0, t1 = float(t[0]), float(t[-1])
        a, b = t0 - overshoot * (t1 - t0), (t1 - t0) * (1.0 + 2.0 * overshoot)
        out.append((a + counter_u01(idx, d, seed) * b).astype(dtype))
    return out


def counter_points_torch(lo, hi, ts, seed, dtype, device, overshoot=0.0, chunk=1 << 26):
    """The same points for the contiguous global range [lo, hi), generated on `device` in chunks."""
    import torch
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    outs = [torch.empty(hi - lo, dtype=tdt, device=device) for _ in ts]
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        idx = torch.arange(c0, c1, dtype=torch.int64, device=device)
        for d, t in enumerate(ts):
            t0, t1 = float(t[0]), float(t[-1])
            a, b = t0 - overshoot * (t1 - t0), (t1 - t0) * (1.0 + 2.0 * overshoot)
            outs[d][c0 - lo:c1 - lo] = (a + counter_u01(idx, d, seed) * b).to(tdt)
    return outs
