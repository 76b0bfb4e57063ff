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


def test_oracle_against_julia_reference_fixtures_when_present():
    """Fixtures written by the UNMODIFIED reference (`julia baseline/bench_reference.jl fixtures`): the oracle must
    reproduce the reference's own values (1e-12 norm-wise) and interval indices (bit-exact).  The directory does not
    exist in this repository — no Julia where it was built — so the test skips; a maintainer with Julia generates
    it and the oracle is then pinned to the reference itself, not only to its KATs and to SciPy."""
    import json
    import os
    import re

    import pytest
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia")
    if not os.path.exists(os.path.join(root, "manifest.json")):
        pytest.skip("no Julia-generated fixtures (tests/golden/julia/manifest.json)")
    arrays = {}
    for e in json.load(open(os.path.join(root, "manifest.json")))["arrays"]:
        dt = {"Float64": np.float64, "Float32": np.float32, "Int64": np.int64}[e["eltype"]]
        arrays[e["name"]] = np.fromfile(os.path.join(root, e["name"] + ".bin"), dtype=dt).reshape(e["shape"], order="F")
    spec = {"lin2": (["linear"] * 2, None), "lin5": (["linear"] * 5, None), "const2": (["constant"] * 2, None),
            "bs3": (["bspline"] * 3, [3, 3, 3]), "bs2d": (["bspline"] * 2, [2, 5]), "nurbs4": (["bspline"] * 4, [2] * 4)}
    checked = 0
    for name, (kinds, degs) in spec.items():
        n = len(kinds)
        for mode in ("unst", "grid"):
            ts = [arrays[f"{name}_{mode}_t{d + 1}"] for d in range(n)]
            xs = [arrays[f"{name}_{mode}_x{d + 1}"] for d in range(n)]
            for d in range(n):
                ref_idx = arrays[f"{name}_{mode}_idx{d + 1}"]
                assert np.array_equal(O.get_idx(kinds[d], ts[d], xs[d], degree=(degs or [0] * n)[d]), ref_idx)
            w = arrays.get(f"{name}_{mode}_w")
            for key, ref in arrays.items():
                m = re.fullmatch(rf"{name}_{mode}_out_(\d+)", key)
                if not m:
                    continue
                orders = [int(c) for c in m.group(1)]
                got = O.evaluate(kinds, ts, arrays[f"{name}_{mode}_u"], xs, degrees=degs, weights=w,
                                 derivative_orders=orders, grid=(mode == "grid"))
                close(got, ref)
                checked += 1
    assert checked >= 20
