"""ctypes front-end of the CPU oracle (oracle/dind_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg.  Never imported by the product package.

Arrays follow the reference's (Julia) layout: column-major.  Pass/receive NumPy arrays in
Fortran order, shape (n_1..n_N, out...) for `u` and (n_points, out...) / (g_1..g_N, out...)
for the result, so that element [i, j, c] here is element [i+1, j+1, c+1] in Julia.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdind_oracle.so")

LINEAR, CONSTANT, BSPLINE = 0, 1, 2
KIND_NAMES = {"linear": LINEAR, "constant": CONSTANT, "bspline": BSPLINE}

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++, -ffp-contract=off, OpenMP)."""
    src = os.path.join(_HERE, "dind_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.dindo_eval.restype = C.c_int
        _lib.dindo_get_idx.restype = C.c_int
        _lib.dindo_basis.restype = C.c_int
        _lib.dindo_expand_knots.restype = C.c_int
        _lib.dindo_max_threads.restype = C.c_int
        _lib.dindo_last_eval_seconds.restype = C.c_double
    return _lib


def last_eval_seconds() -> float:
    """Duration of the last evaluation loop proper (idx/basis cache construction excluded, as in the reference)."""
    return float(lib().dindo_last_eval_seconds())


def max_threads() -> int:
    return int(lib().dindo_max_threads())


def _dt(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return 0
    if dtype == np.float64:
        return 1
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _kind(k):
    return KIND_NAMES[k] if isinstance(k, str) else int(k)


def _vec(a, dtype):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype).ravel())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _mult(m):
    return None if m is None else np.ascontiguousarray(np.asarray(m, dtype=np.int64))


def get_idx(kind, t, t_eval, degree=0, multiplicities=None, dtype=np.float64):
    """idx_eval (1-based Int64, as the reference stores it) — interpolation_utils.jl:116-161."""
    t = _vec(t, dtype)
    te = _vec(t_eval, dtype)
    m = _mult(multiplicities)
    out = np.empty(te.size, dtype=np.int64)
    rc = lib().dindo_get_idx(
        _dt(dtype), _kind(kind), _ptr(t), C.c_int64(t.size), int(degree),
        None if m is None else m.ctypes.data_as(C.c_void_p), _ptr(te), C.c_int64(te.size),
        out.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError(f"oracle get_idx failed rc={rc}")
    return out


def basis(t, degree, t_eval, max_derivative_order_eval=0, multiplicities=None, dtype=np.float64):
    """basis_function_eval cache, shape (n_eval, degree+1, max_deriv+1) — spline_utils.jl:164-191."""
    t = _vec(t, dtype)
    te = _vec(t_eval, dtype)
    m = _mult(multiplicities)
    out = np.empty((te.size, degree + 1, max_derivative_order_eval + 1), dtype=dtype, order="F")
    rc = lib().dindo_basis(
        _dt(dtype), _ptr(t), C.c_int64(t.size), int(degree),
        None if m is None else m.ctypes.data_as(C.c_void_p), _ptr(te), C.c_int64(te.size),
        int(max_derivative_order_eval), _ptr(out))
    if rc:
        raise ValueError(f"oracle basis failed rc={rc}")
    return out


def expand_knots(t, degree, multiplicities=None, dtype=np.float64):
    """knots_all — spline_utils.jl:1-20 with the default clamped multiplicities of
    interpolation_dimensions.jl:168-173."""
    t = _vec(t, dtype)
    m = _mult(multiplicities)
    n = C.c_int64(0)
    mp = None if m is None else m.ctypes.data_as(C.c_void_p)
    lib().dindo_expand_knots(_dt(dtype), _ptr(t), C.c_int64(t.size), int(degree), mp, None, C.byref(n))
    out = np.empty(n.value, dtype=dtype)
    lib().dindo_expand_knots(_dt(dtype), _ptr(t), C.c_int64(t.size), int(degree), mp, _ptr(out), C.byref(n))
    return out


def evaluate(kinds, ts, u, t_evals, *, degrees=None, multiplicities=None, weights=None,
             derivative_orders=None, grid=False, out=None, nthreads=0):
    """eval_unstructured (grid=False) / eval_grid (grid=True) of the reference.

    kinds: per-dimension kind ("linear" | "constant" | "bspline"); ts: per-dimension knots;
    u: array (n_1..n_N, out...), any memory order (copied to Fortran order);
    t_evals: per-dimension evaluation points.  Returns a Fortran-ordered array of
    shape (n_points, out...) or (g_1..g_N, out...), dtype of u.
    If `out` is given (Fortran-ordered, right shape/dtype) it is written in place, so
    "untouched" semantics (Constant derivative, N_out > 0) are observable.
    """
    u = np.asarray(u)
    dtype = u.dtype
    dt = _dt(dtype)
    n_in = len(kinds)
    uf = np.asfortranarray(u)
    kinds_a = (C.c_int * n_in)(*[_kind(k) for k in kinds])
    tv = [_vec(t, dtype) for t in ts]
    tev = [_vec(t, dtype) for t in t_evals]
    t_ptrs = (C.c_void_p * n_in)(*[t.ctypes.data for t in tv])
    te_ptrs = (C.c_void_p * n_in)(*[t.ctypes.data for t in tev])
    n_t = (C.c_int64 * n_in)(*[t.size for t in tv])
    n_eval = (C.c_int64 * n_in)(*[t.size for t in tev])
    degs = (C.c_int * n_in)(*([0] * n_in if degrees is None else [int(d) for d in degrees]))
    mv = [None] * n_in if multiplicities is None else [_mult(m) for m in multiplicities]
    m_ptrs = (C.c_void_p * n_in)(*[None if m is None else m.ctypes.data for m in mv])
    orders = (C.c_int * n_in)(*([0] * n_in if derivative_orders is None else [int(o) for o in derivative_orders]))
    u_dims = (C.c_int64 * uf.ndim)(*uf.shape)
    wf = None if weights is None else np.asfortranarray(np.asarray(weights, dtype=dtype))
    out_shape = (tuple(t.size for t in tev) if grid else (tev[0].size,)) + tuple(uf.shape[n_in:])
    if out is None:
        out = np.zeros(out_shape, dtype=dtype, order="F")
    else:
        assert out.shape == out_shape and out.dtype == dtype and out.flags.f_contiguous
    rc = lib().dindo_eval(
        dt, n_in, kinds_a, t_ptrs, n_t, degs, m_ptrs, _ptr(uf), u_dims, uf.ndim,
        None if wf is None else _ptr(wf), te_ptrs, n_eval, orders, int(bool(grid)), _ptr(out),
        int(nthreads))
    if rc:
        raise ValueError(f"oracle evaluate failed rc={rc}")
    return out


def point(kinds, ts, u, t, **kw):
    """Single-point call interp(t...) (src/DataInterpolationsND.jl:85-100): a 1-point
    unstructured evaluation (the cached and on-the-fly basis values are the same numbers)."""
    res = evaluate(kinds, ts, u, [[x] for x in t], grid=False, **kw)
    return res[0]
