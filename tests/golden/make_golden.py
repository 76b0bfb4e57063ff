"""Generates tests/golden/golden_scipy.npz — INDEPENDENT golden vectors for the hot path.

The reference (pure Julia) cannot run in this image, so these vectors come from SciPy's
own B-spline / regular-grid implementations (scipy.interpolate.BSpline and
RegularGridInterpolator), i.e. from an implementation that shares no code with either the
oracle (oracle/dind_oracle.cpp) or the CUDA library.  The oracle is pinned against them in
tests/test_oracle_golden.py, the CUDA path in tests/test_gpu_parity.py.

Run:  python tests/golden/make_golden.py     (needs scipy; deterministic, seed fixed)
"""
import os

import numpy as np
from scipy.interpolate import BSpline, RegularGridInterpolator

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(20261019)


def expand(t, degree, mult=None):
    t = np.asarray(t, dtype=np.float64)
    if mult is None:
        mult = np.ones(t.size, dtype=np.int64)
        mult[0] = mult[-1] = degree + 1
    return np.repeat(t, mult)


def basis_matrix(knots_all, degree, x, r):
    """M[i, j] = d^r/dx^r B_j(x_i); polynomial extrapolation outside the knot span
    (what the reference's clamped interval index gives)."""
    n = knots_all.size - degree - 1
    M = np.zeros((x.size, n))
    for j in range(n):
        c = np.zeros(n)
        c[j] = 1.0
        b = BSpline(knots_all, c, degree, extrapolate=True)
        if r > degree:
            continue
        M[:, j] = (b.derivative(r) if r > 0 else b)(x)
    return M


def query_points(t, n_rand, outside=True):
    t = np.asarray(t, dtype=np.float64)
    pts = [t[:-1] + np.diff(t) / 2, rng.uniform(t[0], t[-1], n_rand)]
    if outside:
        span = t[-1] - t[0]
        pts.append(np.array([t[0] - 0.13 * span, t[-1] + 0.21 * span]))
    return np.concatenate(pts)


out = {}

# --- 1-D B-splines, degrees 1..5, default clamped multiplicities, derivative orders 0..3 ----
t1d = np.array([1.0, 2.0, 3.5, 4.0, 10.0, 11.5, 13.0])
for p in (1, 2, 3, 4, 5):
    ka = expand(t1d, p)
    n = ka.size - p - 1
    u = rng.uniform(-1, 1, n)
    x = query_points(t1d, 12)
    out[f"bs1d_p{p}_t"] = t1d
    out[f"bs1d_p{p}_u"] = u
    out[f"bs1d_p{p}_x"] = x
    for r in range(0, 4):
        out[f"bs1d_p{p}_r{r}"] = basis_matrix(ka, p, x, r) @ u

# --- 1-D B-spline with interior multiplicities -------------------------------------------
tm = np.array([0.0, 0.7, 1.1, 2.9, 4.0])
mult = np.array([3, 1, 2, 1, 3])
ka = expand(tm, 2, mult)
n = ka.size - 2 - 1
u = rng.uniform(-1, 1, n)
x = np.concatenate([tm[:-1] + np.diff(tm) / 2, rng.uniform(tm[0], tm[-1], 10)])
out["bsmult_t"], out["bsmult_m"], out["bsmult_u"], out["bsmult_x"] = tm, mult, u, x
out["bsmult_r0"] = basis_matrix(ka, 2, x, 0) @ u
out["bsmult_r1"] = basis_matrix(ka, 2, x, 1) @ u

# --- 2-D tensor-product B-spline, degrees (2, 3), vector valued (2 components) ------------
ta = np.array([-3.14, 1.0, 3.0, 7.6, 12.8])
tb = np.array([-2.71, 1.41, 12.76, 50.2, 120.0])
ka_a, ka_b = expand(ta, 2), expand(tb, 3)
na, nb = ka_a.size - 3, ka_b.size - 4
u2 = rng.uniform(-1, 1, (na, nb, 2))
xa, xb = query_points(ta, 9), query_points(tb, 9)
out["bs2d_ta"], out["bs2d_tb"], out["bs2d_u"], out["bs2d_xa"], out["bs2d_xb"] = ta, tb, u2, xa, xb
for ra, rb in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 1)):
    Ma, Mb = basis_matrix(ka_a, 2, xa, ra), basis_matrix(ka_b, 3, xb, rb)
    out[f"bs2d_grid_r{ra}{rb}"] = np.einsum("ia,jb,abc->ijc", Ma, Mb, u2)     # eval_grid
    out[f"bs2d_unst_r{ra}{rb}"] = np.einsum("ia,ib,abc->ic", Ma, Mb, u2)      # eval_unstructured (zip)

# --- 3-D cubic B-spline, scalar, unstructured --------------------------------------------
t3 = [np.cumsum(0.5 + rng.uniform(0, 1, 6)) for _ in range(3)]
ka3 = [expand(t, 3) for t in t3]
n3 = [k.size - 4 for k in ka3]
u3 = rng.uniform(0, 1, n3)
x3 = [rng.uniform(t[0], t[-1], 40) for t in t3]
for d in range(3):
    out[f"bs3d_t{d}"], out[f"bs3d_x{d}"] = t3[d], x3[d]
out["bs3d_u"] = u3
for orders in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)):
    M = [basis_matrix(ka3[d], 3, x3[d], orders[d]) for d in range(3)]
    out["bs3d_r%d%d%d" % orders] = np.einsum("ia,ib,ic,abc->i", M[0], M[1], M[2], u3)

# --- N-D linear (RegularGridInterpolator, linear extrapolation outside) --------------------
for nd, shape in ((2, (6, 5)), (3, (5, 4, 6)), (5, (3, 4, 3, 4, 3))):
    ts = [np.cumsum(0.5 + rng.uniform(0, 1, n)) for n in shape]
    u = rng.uniform(-1, 1, shape)
    rgi = RegularGridInterpolator(ts, u, method="linear", bounds_error=False, fill_value=None)
    npt = 50
    xs = [rng.uniform(t[0] - 0.2, t[-1] + 0.2, npt) for t in ts]
    for d in range(nd):
        out[f"lin{nd}d_t{d}"], out[f"lin{nd}d_x{d}"] = ts[d], xs[d]
    out[f"lin{nd}d_u"] = u
    out[f"lin{nd}d_val"] = rgi(np.stack(xs, axis=1))

np.savez_compressed(os.path.join(HERE, "golden_scipy.npz"), **out)
print("wrote", os.path.join(HERE, "golden_scipy.npz"), len(out), "arrays")
