"""Randomised cross-check of the specialised kernels against the generic device kernels and the CPU
oracle (pytest -m gpu): random dimension counts, kinds, degrees, grid sizes (odd sizes, sizes below
one warp, unsorted / repeated / sparse t_eval, points outside the knot span, several components)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
GENERIC = 8


@pytest.fixture(scope="module")
def D():
    import dind_b200
    assert dind_b200._lib.load().dind_device_count() > 0
    return dind_b200


def _random_case(rng, grid):
    n_in = int(rng.integers(2, 5)) if grid else int(rng.integers(1, 5))
    mode = rng.choice(["linear", "bspline", "bspline_const"])
    kinds, degrees = [], []
    p = int(rng.integers(1, 4))
    for d in range(n_in):
        if mode == "linear":
            kinds.append("linear"); degrees.append(0)
        elif mode == "bspline_const" and grid and 2 <= d < n_in - 1 and rng.random() < 0.7:
            kinds.append("constant"); degrees.append(0)
        else:
            kinds.append("bspline"); degrees.append(p)
    dtype = rng.choice([np.float32, np.float64])
    n_knots = [int(rng.integers(4, 12)) for _ in range(n_in)]
    ts = [H.knots(rng, n, dtype, uniform=bool(rng.integers(0, 2))) for n in n_knots]
    n_u = [t.size + q - 1 if k == "bspline" else t.size for k, q, t in zip(kinds, degrees, ts)]
    out_dims = tuple(int(x) for x in rng.integers(1, 4, size=int(rng.integers(0, 2))))
    u = rng.random(tuple(n_u) + out_dims).astype(dtype)
    return kinds, degrees, ts, u, dtype


def _t_eval(rng, t, n, dtype):
    style = rng.choice(["sorted_dense", "sorted_sparse", "unsorted", "repeated"])
    lo, hi = float(t[0]), float(t[-1])
    span = hi - lo
    if style == "sorted_dense":
        x = np.sort(rng.uniform(lo - 0.1 * span, hi + 0.1 * span, n))
    elif style == "sorted_sparse":
        x = np.sort(rng.choice(np.linspace(lo, hi, 7), n))
    elif style == "unsorted":
        x = rng.uniform(lo, hi, n)
    else:
        x = np.repeat(rng.uniform(lo, hi, max(1, n // 3)), 3)[:n]
        if x.size < n:
            x = np.concatenate([x, np.full(n - x.size, lo)])
    return x.astype(dtype)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_grid_fast_vs_generic_vs_oracle(D, seed):
    rng = np.random.default_rng(9000 + seed)
    kinds, degrees, ts, u, dtype = _random_case(rng, grid=True)
    sizes = [int(rng.choice([1, 3, 31, 33, 64, 70, 130])) if d == 0 else int(rng.integers(1, 12)) for d in range(len(kinds))]
    if np.prod(sizes) > 400_000:
        sizes[0] = 33
    xs = [_t_eval(rng, t, n, dtype) for t, n in zip(ts, sizes)]
    orders = tuple(int(rng.integers(0, 2)) if k != "constant" else 0 for k in kinds)
    kw = dict(degrees=degrees, derivative_orders=orders, grid=True)
    ref = H.oracle_eval(kinds, ts, u, xs, **kw)
    fast, kern = H.gpu_eval(D, kinds, ts, u, xs, **kw)
    gen, _ = H.gpu_eval(D, kinds, ts, u, xs, flags=GENERIC, **kw)
    H.assert_close(fast, ref, dtype)
    H.assert_close(gen, ref, dtype)


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_unstructured_fast_vs_generic_vs_oracle(D, seed):
    rng = np.random.default_rng(7000 + seed)
    kinds, degrees, ts, u, dtype = _random_case(rng, grid=False)
    n = int(rng.choice([1, 31, 257, 5000]))
    xs = [_t_eval(rng, t, n, dtype) for t in ts]
    orders = tuple(int(rng.integers(0, 2)) for _ in kinds)
    kw = dict(degrees=degrees, derivative_orders=orders)
    ref = H.oracle_eval(kinds, ts, u, xs, **kw)
    for flags in (0, GENERIC, 16):
        got, _ = H.gpu_eval(D, kinds, ts, u, xs, flags=flags, **kw)
        H.assert_close(got, ref, dtype)
