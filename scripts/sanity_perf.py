#!/usr/bin/env python
"""Quick throughput sanity check of common configurations outside the five BASELINE configs
(run on a B200): prints Gpoints/s and algorithmic GB/s per case, to catch pathologies (spills,
fallbacks to the generic kernels)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dind_b200 as D  # noqa: E402


def knots(rng, n, dtype):
    return np.cumsum(0.5 + rng.random(n)).astype(dtype)


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""   # substring filter: python scripts/sanity_perf.py "tricubic F32 256"


def run(name, kinds, degrees, n_u, n_grid, dtype, nurbs=False, out_dims=()):
    if ONLY and ONLY not in name:
        return
    rng = np.random.default_rng(0)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    ts, dims = [], []
    for k, p, n, g in zip(kinds, degrees, n_u, n_grid):
        t = knots(rng, n - p + 1 if k == "bspline" else n, dtype)
        x = torch.from_numpy(np.linspace(t[0], t[-1], g).astype(dtype)).cuda()
        ts.append(t)
        dims.append(D.LinearInterpolationDimension(t, t_eval=x) if k == "linear" else
                    D.ConstantInterpolationDimension(t, t_eval=x) if k == "constant" else
                    D.BSplineInterpolationDimension(t, p, t_eval=x, max_derivative_order_eval=1))
    shape = tuple(n_u) + tuple(out_dims)
    u = D.f_empty(shape, dtype, "cuda")
    u.copy_(torch.rand(shape, device="cuda", dtype=tdt))
    cache = None
    if nurbs:
        w = D.f_empty(tuple(n_u), dtype, "cuda")
        w.copy_(torch.rand(tuple(n_u), device="cuda", dtype=tdt) + 0.5)
        cache = D.NURBSWeights(w)
    itp = D.NDInterpolation(u, tuple(dims), cache=cache, stream=torch.cuda.current_stream().cuda_stream)
    out = D.f_empty(tuple(n_grid) + tuple(out_dims), dtype, "cuda")
    for _ in range(3):
        D.eval_grid_(out, itp, flags=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        D.eval_grid_(out, itp, flags=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    pts = float(np.prod(n_grid))
    esz = np.dtype(dtype).itemsize
    nbytes = (pts * max(1, int(np.prod(out_dims))) + float(np.prod(shape)) * (2 if nurbs else 1)) * esz
    print(f"{name:46s} {itp.last_kernel:36s} {pts / ms / 1e6:9.1f} Gpt/s {nbytes / ms / 1e6:8.0f} GB/s  {ms:8.3f} ms")
    itp.close()


if __name__ == "__main__":
    f32, f64 = np.float32, np.float64
    run("2-D bilinear F32 4096^2 -> 8192^2", ["linear"] * 2, [0, 0], (4096, 4096), (8192, 8192), f32)
    run("2-D bicubic F64 1024^2 -> 4096^2", ["bspline"] * 2, [3, 3], (1024, 1024), (4096, 4096), f64)
    run("3-D trilinear F64 256^3 -> 512^3", ["linear"] * 3, [0] * 3, (256,) * 3, (512,) * 3, f64)
    run("3-D tricubic F32 128^3 -> 384^3", ["bspline"] * 3, [3] * 3, (128,) * 3, (384,) * 3, f32)
    run("3-D tricubic F64 128^3 -> 384^3", ["bspline"] * 3, [3] * 3, (128,) * 3, (384,) * 3, f64)
    run("3-D tricubic F64 128^3 -> 256^3, 3 comps (config-3 u)", ["bspline"] * 3, [3] * 3, (128,) * 3, (256,) * 3, f64, out_dims=(3,))
    run("3-D tricubic F32 256^3 -> 512^3", ["bspline"] * 3, [3] * 3, (256,) * 3, (512,) * 3, f32)
    run("3-D triquadratic F64 128^3 -> 256^3, 3 comps", ["bspline"] * 3, [2] * 3, (128,) * 3, (256,) * 3, f64, out_dims=(3,))
    run("3-D trilinear F32 256^3 -> 500x333x250 (odd)", ["linear"] * 3, [0] * 3, (256,) * 3, (500, 333, 250), f32)
    run("4-D linear F32 64^4 -> 128^4", ["linear"] * 4, [0] * 4, (64,) * 4, (128,) * 4, f32)
    run("3-D const x lin x lin F32 (generic path)", ["constant", "linear", "linear"], [0] * 3, (128,) * 3, (256,) * 3, f32)
    run("3-D all-Constant F32 256^3 -> 512^3", ["constant"] * 3, [0] * 3, (256,) * 3, (512,) * 3, f32)
