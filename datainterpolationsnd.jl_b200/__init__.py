"""Host-side mirror of DataInterpolationsND.jl's API for the batched evaluation hot path, on
top of the C ABI of libdind.so (include/dind.h).

The reference's host language is Julia, which this image does not have; the Julia package
that `ccall`s the same ABI ships as source under julia/ (unexecuted).  This Python package
binds the identical symbols with ctypes and keeps the reference's names, argument meaning
and error behaviour (src/DataInterpolationsND.jl:102-104 exports):

    LinearInterpolationDimension(t; t_eval)                 src/interpolation_dimensions.jl:37
    ConstantInterpolationDimension(t; left, t_eval)         src/interpolation_dimensions.jl:87
    BSplineInterpolationDimension(t, degree; t_eval, max_derivative_order_eval, multiplicities)  :162
    NURBSWeights(weights)                                   src/spline_utils.jl:232
    NDInterpolation(u, interp_dims; cache)                  src/DataInterpolationsND.jl:29-56
    eval_unstructured / eval_unstructured_ (Julia: eval_unstructured!)   src/interpolation_parallel.jl:14-52
    eval_grid / eval_grid_ (Julia: eval_grid!)              src/interpolation_parallel.jl:66-103
    interp(t...) / interp(t..., out=out)                    src/DataInterpolationsND.jl:73-100

Arrays are column-major like Julia's: NumPy arrays are taken in / returned in Fortran order
with the reference's shapes; CUDA torch tensors are accepted zero-copy when their strides
are column-major (see `f_empty`).  All evaluation happens on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DindError, shard_range  # noqa: F401
from .parallel import eval_unstructured_allgather_, gather_grid, gather_unstructured, shard_t_eval  # noqa: F401

__all__ = [
    "NDInterpolation", "LinearInterpolationDimension", "ConstantInterpolationDimension",
    "BSplineInterpolationDimension", "NURBSWeights", "EmptyCache", "eval_unstructured",
    "eval_unstructured_", "eval_grid", "eval_grid_", "DindError", "f_empty", "shard_range",
    "shard_t_eval", "gather_unstructured", "gather_grid", "eval_unstructured_allgather_", "eval_unstructured_gradient",
]


# ---- array helpers ----------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _np_dtype(x):
    if _is_torch(x):
        import torch
        return {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64)}.get(x.dtype, np.dtype(np.float64))
    return np.asarray(x).dtype


def _is_f_contiguous_torch(x):
    return x.permute(*reversed(range(x.ndim))).is_contiguous() if x.ndim > 1 else x.is_contiguous()


def f_empty(shape, dtype, device):
    """A torch tensor of `shape` whose memory is column-major (what libdind reads/writes)."""
    import torch
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
    shape = tuple(int(s) for s in shape)
    base = torch.empty(shape[::-1], dtype=tdt, device=device)
    return base.permute(*reversed(range(len(shape)))) if len(shape) > 1 else base


def _compute_dtype(arrays):
    """promote_type over element types: Float32 only if everything is Float32."""
    dts = [_np_dtype(a) for a in arrays if a is not None]
    return np.dtype(np.float32) if dts and all(d == np.float32 for d in dts) else np.dtype(np.float64)


class _Arr:
    """A pointer + keep-alive for an array handed to the C ABI in the compute dtype."""

    def __init__(self, x, dtype, fortran=True):
        self.shape = tuple(x.shape)
        if _is_torch(x) and x.is_cuda:
            import torch
            tdt = torch.float32 if dtype == np.float32 else torch.float64
            if x.dtype != tdt:
                x = x.to(tdt)
            if not _is_f_contiguous_torch(x):
                raise DindError(_lib.ERR_INVALID_ARGUMENT,
                                "CUDA tensors must be column-major (see f_empty); got strides %s" % (x.stride(),))
            self.keep, self.ptr, self.device = x, x.data_ptr(), True
        else:
            if _is_torch(x):
                x = x.cpu().numpy()
            a = np.asarray(x, dtype=dtype)
            a = np.asfortranarray(a) if fortran else np.ascontiguousarray(a)
            self.keep, self.ptr, self.device = a, a.ctypes.data, False


# ---- interpolation dimensions -------------------------------------------------------------------
class AbstractInterpolationDimension:
    kind = None
    degree = 0
    multiplicities = None
    max_derivative_order_eval = 0
    left = True

    def _init_common(self, t, t_eval):
        self.t = np.asarray(t) if not _is_torch(t) else t.cpu().numpy()
        if self.t.ndim != 1:
            raise DindError(_lib.ERR_INVALID_ARGUMENT, "t must be an AbstractVector with number like elements.")
        if not np.issubdtype(self.t.dtype, np.floating):
            self.t = self.t.astype(np.float64)
        if self.t.size > 1 and not np.all(np.diff(self.t) > 0):   # validate_t (interpolation_utils.jl:34-37)
            raise DindError(_lib.ERR_INVALID_ARGUMENT, "The elements of t must be sorted and unique.")
        self.set_t_eval(t_eval)

    def set_t_eval(self, t_eval):
        if t_eval is None:
            t_eval = np.empty(0, dtype=self.t.dtype)   # similar(t, 0)
        elif not _is_torch(t_eval):
            t_eval = np.asarray(t_eval)
            if not np.issubdtype(t_eval.dtype, np.floating):
                t_eval = t_eval.astype(np.float64)
        self.t_eval = t_eval
        self._cache = {}

    def __len__(self):   # Base.length(itp_dim) = length(itp_dim.t) (interpolation_utils.jl:3)
        return int(self.t.size)

    def _dtype(self):
        return _compute_dtype([self.t, self.t_eval])

    def _solo(self):
        """A private 1-D handle used for the cache fields (idx_eval, basis_function_eval)."""
        h = _Handle(1, self._dtype(), None)
        h.set_dim(0, self)
        h.set_t_eval(0, self)
        return h

    @property
    def idx_eval(self):
        """idx_eval::Vector{Int}, 1-based (set_eval_idx!, interpolation_utils.jl:143-161), computed on the GPU."""
        if "idx" not in self._cache:
            n = int(self.t_eval.shape[0])
            out = np.empty(n, dtype=np.int64)
            if n:
                h = self._solo()
                _lib.check(_lib.load().dind_get_idx_eval(h.h, 0, out.ctypes.data_as(C.POINTER(C.c_int64))))
                h.close()
            self._cache["idx"] = out
        return self._cache["idx"]

    def n_u(self):
        return int(self.t.size)


class LinearInterpolationDimension(AbstractInterpolationDimension):
    kind = _lib.LINEAR

    def __init__(self, t, *, t_eval=None):
        self._init_common(t, t_eval)


class ConstantInterpolationDimension(AbstractInterpolationDimension):
    kind = _lib.CONSTANT

    def __init__(self, t, *, left=True, t_eval=None):
        self._init_common(t, t_eval)
        self.left = bool(left)   # stored, never read by evaluation — like the reference's field


class BSplineInterpolationDimension(AbstractInterpolationDimension):
    kind = _lib.BSPLINE

    def __init__(self, t, degree, *, t_eval=None, max_derivative_order_eval=0, multiplicities=None):
        self._init_common(t, t_eval)
        self.degree = int(degree)
        self.max_derivative_order_eval = int(max_derivative_order_eval)
        if multiplicities is None:   # open/clamped knot vector (interpolation_dimensions.jl:168-173)
            multiplicities = np.ones(self.t.size, dtype=np.int64)
            multiplicities[[0, -1]] = self.degree + 1
        self.multiplicities = np.asarray(multiplicities, dtype=np.int64)
        if self.multiplicities.size != self.t.size:
            raise DindError(_lib.ERR_SIZE_MISMATCH, "There must be the same amount of knots (points in t) as multiplicities.")
        if not np.all((1 <= self.multiplicities) & (self.multiplicities <= self.degree + 1)):
            raise DindError(_lib.ERR_INVALID_ARGUMENT, "All multiplicities must be between 1 and degree + 1.")
        self.knots_all = np.repeat(self.t, self.multiplicities)   # expand_knot_vector_kernel (spline_utils.jl:1-20)

    def n_u(self):   # get_n_basis_functions (spline_utils.jl:223-225)
        return int(self.knots_all.size - self.degree - 1)

    @property
    def basis_function_eval(self):
        """basis_function_eval[n_eval, degree+1, max_derivative_order_eval+1] (spline_utils.jl:164-191), from the GPU."""
        if "basis" not in self._cache:
            n = int(self.t_eval.shape[0])
            out = np.empty((n, self.degree + 1, self.max_derivative_order_eval + 1), dtype=self._dtype(), order="F")
            if n:
                h = self._solo()
                _lib.check(_lib.load().dind_get_basis_function_eval(h.h, 0, out.ctypes.data_as(C.c_void_p)))
                h.close()
            self._cache["basis"] = out
        return self._cache["basis"]


class EmptyCache:
    pass


class NURBSWeights:
    def __init__(self, weights):
        self.weights = weights


# ---- thin handle wrapper ---------------------------------------------------------------------------
class _Handle:
    def __init__(self, n_in, dtype, device):
        self.lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.h = C.c_void_p()
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except ImportError:
                pass
        self.device = int(device)
        _lib.check(self.lib.dind_create(C.byref(self.h), int(n_in), _lib.F32 if self.dtype == np.float32 else _lib.F64, self.device))
        self.keep = {}

    def close(self):
        if self.h:
            self.lib.dind_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_dim(self, d, dim):
        t = np.ascontiguousarray(dim.t, dtype=self.dtype)
        m = None
        if dim.kind == _lib.BSPLINE:
            m = np.ascontiguousarray(dim.multiplicities, dtype=np.int64).ctypes.data_as(C.POINTER(C.c_int64))
        _lib.check(self.lib.dind_set_dim(self.h, d, dim.kind, t.ctypes.data, t.size, int(dim.degree), m, int(dim.left)))

    def set_t_eval(self, d, dim):
        a = _Arr(dim.t_eval, self.dtype)
        self.keep[("t_eval", d)] = a
        _lib.check(self.lib.dind_set_t_eval(self.h, d, a.ptr if a.shape[0] else None, a.shape[0],
                                            int(dim.max_derivative_order_eval), 0))

    def set_stream(self, stream_ptr):
        _lib.check(self.lib.dind_set_stream(self.h, C.c_void_p(stream_ptr)))


def _orders(derivative_orders, n_in):
    if derivative_orders is None:
        return None
    derivative_orders = tuple(int(o) for o in derivative_orders)
    if len(derivative_orders) != n_in:
        raise DindError(_lib.ERR_SIZE_MISMATCH, f"derivative_orders must have {n_in} entries")
    return (C.c_int * n_in)(*derivative_orders)


class NDInterpolation:
    """NDInterpolation(u, interp_dims; cache = EmptyCache()) — src/DataInterpolationsND.jl:29-56."""

    def __init__(self, u, interp_dims, cache=None, *, device=None, stream=None):
        if isinstance(interp_dims, AbstractInterpolationDimension):
            interp_dims = (interp_dims,)
        self.interp_dims = tuple(interp_dims)
        self.cache = EmptyCache() if cache is None else cache
        self.N_in = len(self.interp_dims)
        if not _is_torch(u):
            u = np.asarray(u)
            if not np.issubdtype(u.dtype, np.floating):
                u = u.astype(np.float64)
        self.N_out = u.ndim - self.N_in
        if self.N_out < 0:
            raise DindError(_lib.ERR_SIZE_MISMATCH,
                            "The number of dimensions of u must be at least the number of interpolation dimensions.")
        weights = self.cache.weights if isinstance(self.cache, NURBSWeights) else None
        arrays = [u, weights] + [d.t for d in self.interp_dims] + [d.t_eval for d in self.interp_dims]
        self.dtype = _compute_dtype(arrays)
        self.out_dtype = _np_dtype(u)          # similar(u, ...) keeps eltype(u)
        self._h = _Handle(self.N_in, self.dtype, device)
        if stream is not None:
            self._h.set_stream(stream)
        for d, dim in enumerate(self.interp_dims):
            self._h.set_dim(d, dim)
            self._h.set_t_eval(d, dim)
        if weights is not None:
            self._set_weights(weights)
        self.set_u(u)

    # -- data ------------------------------------------------------------------------------
    def set_t_eval(self, t_evals):
        """Swap in new evaluation points for every dimension (the reference rebuilds the dimension
        structs for that: `t_eval` is a constructor keyword).  All caches that depend on t_eval are
        dropped and rebuilt lazily, or explicitly by `build_caches`."""
        t_evals = tuple(t_evals)
        if len(t_evals) != self.N_in:
            raise DindError(_lib.ERR_SIZE_MISMATCH, f"expected {self.N_in} t_eval arrays")
        for d, (dim, x) in enumerate(zip(self.interp_dims, t_evals)):
            dim.set_t_eval(x)
            self._h.set_t_eval(d, dim)
        return self

    def set_u(self, u, flags=0):
        """Swap in a new `u` of the same shape (re-evaluation with the t_eval caches kept).
        flags=1 (DIND_FLAG_ASYNC) with a pinned host array: the upload is only queued — keep `u`
        unchanged until `synchronize()`.

        A CUDA tensor is BORROWED (zero-copy), and the library keeps copies derived from it (the
        component-fastest `u` of vector-valued fields and its cell-record form for large multilinear tables, `u * weights` for NURBS): after changing a device
        `u` (or the weights) IN PLACE, call `set_u` again before evaluating."""
        a = _Arr(u, self.dtype)
        if len(a.shape) - self.N_in != self.N_out:
            raise DindError(_lib.ERR_SIZE_MISMATCH, "u changed its number of dimensions")
        dims = (C.c_int64 * len(a.shape))(*a.shape)
        _lib.check(self._h.lib.dind_set_u(self._h.h, C.c_void_p(a.ptr), dims, len(a.shape), int(flags)))
        self._h.keep["u"] = a
        self.u = u
        self._u_on_device = a.device
        return self

    def _set_weights(self, w):
        a = _Arr(w, self.dtype)
        dims = (C.c_int64 * len(a.shape))(*a.shape)
        _lib.check(self._h.lib.dind_set_weights(self._h.h, C.c_void_p(a.ptr), dims, len(a.shape), 0))
        self._h.keep["w"] = a

    def set_stream(self, stream_ptr):
        self._h.set_stream(stream_ptr)

    def synchronize(self):
        """Wait for everything queued with DIND_FLAG_ASYNC (kernels and host copies)."""
        _lib.check(self._h.lib.dind_synchronize(self._h.h))

    @property
    def output_size(self):   # get_output_size (interpolation_utils.jl:77-79)
        return tuple(self._h.keep["u"].shape[self.N_in:])

    @property
    def last_kernel(self):
        return self._h.lib.dind_last_kernel(self._h.h).decode()

    def plan_permutation(self):
        """perm (uint32, n_points) of the cell-sorted plan: plan position k holds original point perm[k]."""
        n = int(self.interp_dims[0].t_eval.shape[0])
        out = np.empty(n, dtype=np.uint32)
        _lib.check(self._h.lib.dind_get_plan_permutation(self._h.h, out.ctypes.data_as(C.c_void_p)))
        return out

    @property
    def launch_count(self):
        return int(self._h.lib.dind_launch_count(self._h.h))

    def close(self):
        self._h.close()

    def build_caches(self, *, grid=False, derivative_orders=None, flags=0):
        """Build everything that depends on t_eval / u but not on the evaluation itself (axis tables of
        eval_grid; component-fastest u, idx cache with FLAG_CACHED_IDX, cell-sorted plan with
        FLAG_CELL_SORTED for eval_unstructured) — the constructors' set_eval_idx! / set_basis_function_eval!
        step (src/interpolation_utils.jl:143-161, src/spline_utils.jl:164-191) as an explicit, timeable call."""
        _lib.check(self._h.lib.dind_build_caches(self._h.h, 1 if grid else 0, _orders(derivative_orders, self.N_in), int(flags)))

    # -- evaluation --------------------------------------------------------------------------
    def _make_out(self, shape):
        if self._u_on_device:
            return f_empty(shape, self.dtype, f"cuda:{self._h.device}")
        return np.zeros(shape, dtype=self.dtype, order="F")

    def _out_ptr(self, out, shape):
        """Validate a user-provided `out`; returns (pointer, finalize) where finalize copies back if a cast was needed."""
        if tuple(out.shape) != tuple(shape):
            raise DindError(_lib.ERR_SIZE_MISMATCH, f"out has shape {tuple(out.shape)}, expected {tuple(shape)}")
        if _is_torch(out) and out.is_cuda:
            if _np_dtype(out) != self.dtype or not _is_f_contiguous_torch(out):
                raise DindError(_lib.ERR_INVALID_ARGUMENT, "device `out` must be column-major and of the compute dtype")
            return out.data_ptr(), lambda: out
        if isinstance(out, np.ndarray) and out.dtype == self.dtype and (out.flags.f_contiguous or out.ndim <= 1):
            return out.ctypes.data, lambda: out
        tmp = np.asfortranarray(np.asarray(out, dtype=self.dtype))

        def fin():
            out[...] = tmp
            return out
        return tmp.ctypes.data, fin, tmp

    def _eval(self, fn, out, shape, derivative_orders, flags):
        if out is None:
            out_arr = self._make_out(shape)
            res = self._out_ptr(out_arr, shape)
        else:
            res = self._out_ptr(out, shape)
        _lib.check(fn(self._h.h, C.c_void_p(res[0]), _orders(derivative_orders, self.N_in), int(flags)))
        result = res[1]()
        if out is None and not _is_torch(result) and result.dtype != self.out_dtype:
            result = result.astype(self.out_dtype)   # similar(u, ...): eltype of u
        return result

    def __call__(self, *t, out=None, derivative_orders=None):
        """interp(t...) / interp(out, t...) — src/DataInterpolationsND.jl:73-100 (single point, uncached)."""
        if len(t) == 1 and isinstance(t[0], (tuple, list)):
            t = tuple(t[0])
        if len(t) != self.N_in:
            raise DindError(_lib.ERR_SIZE_MISMATCH, f"expected {self.N_in} coordinates, got {len(t)}")
        pts = [np.array([x], dtype=self.dtype) for x in t]
        ptrs = (C.c_void_p * self.N_in)(*[p.ctypes.data for p in pts])
        shape = (1,) + self.output_size
        buf = np.zeros(shape, dtype=self.dtype, order="F")
        if out is not None:
            if tuple(np.shape(out)) != self.output_size:   # :92
                raise DindError(_lib.ERR_SIZE_MISMATCH, "The size of out must match the size of the last N_out dimensions of u.")
            buf[0] = out   # untouched semantics of the Constant derivative
        _lib.check(self._h.lib.dind_eval_points(self._h.h, buf.ctypes.data, ptrs, 1, _orders(derivative_orders, self.N_in), 0))
        if out is not None:
            out[...] = buf[0]
            return out
        return buf[0, ...][()] if self.N_out == 0 else np.array(buf[0])

    def eval_points(self, t, out=None, derivative_orders=None, flags=0):
        """Batched uncached evaluation at zipped points t = (t_1, ..., t_N) (arrays of equal length)."""
        arrs = [_Arr(x, self.dtype) for x in t]
        if len(arrs) != self.N_in or any(a.shape != arrs[0].shape or len(a.shape) != 1 for a in arrs):
            raise DindError(_lib.ERR_SIZE_MISMATCH, "t must be N_in vectors of equal length")
        if len({a.device for a in arrs}) != 1:
            raise DindError(_lib.ERR_INVALID_ARGUMENT, "all t must be on the host or all on the device")
        n = arrs[0].shape[0]
        ptrs = (C.c_void_p * self.N_in)(*[a.ptr for a in arrs])
        shape = (n,) + self.output_size
        if out is None:
            out = f_empty(shape, self.dtype, f"cuda:{self._h.device}") if arrs[0].device else np.zeros(shape, dtype=self.dtype, order="F")
        res = self._out_ptr(out, shape)
        _lib.check(self._h.lib.dind_eval_points(self._h.h, C.c_void_p(res[0]), ptrs, n, _orders(derivative_orders, self.N_in), int(flags)))
        return res[1]()


def eval_unstructured(interp, *, derivative_orders=None, flags=0):
    """eval_unstructured(interp; derivative_orders) — src/interpolation_parallel.jl:14-18."""
    n_points = int(interp.interp_dims[0].t_eval.shape[0])
    return interp._eval(interp._h.lib.dind_eval_unstructured, None, (n_points,) + interp.output_size, derivative_orders, flags)


def eval_unstructured_(out, interp, *, derivative_orders=None, flags=0):
    """eval_unstructured!(out, interp; derivative_orders) — src/interpolation_parallel.jl:34-52."""
    n_points = int(interp.interp_dims[0].t_eval.shape[0])
    return interp._eval(interp._h.lib.dind_eval_unstructured, out, (n_points,) + interp.output_size, derivative_orders, flags)


def eval_unstructured_gradient(interp, out=None, *, flags=0):
    """Value and all first partial derivatives at the cached unstructured points in one pass
    (dind_eval_unstructured_gradient; beyond the reference, which needs N_in + 1 eval_unstructured calls).
    Returns / fills an array of shape (n_points, out..., N_in + 1): [..., 0] = value, [..., d] = d/dt_d."""
    n_points = int(interp.interp_dims[0].t_eval.shape[0])
    shape = (n_points,) + interp.output_size + (interp.N_in + 1,)
    out_arr = interp._make_out(shape) if out is None else out
    res = interp._out_ptr(out_arr, shape)
    _lib.check(interp._h.lib.dind_eval_unstructured_gradient(interp._h.h, C.c_void_p(res[0]), int(flags)))
    return res[1]()


def eval_grid(interp, *, derivative_orders=None, flags=0):
    """eval_grid(interp; derivative_orders) — src/interpolation_parallel.jl:66-70."""
    grid = tuple(int(d.t_eval.shape[0]) for d in interp.interp_dims)
    return interp._eval(interp._h.lib.dind_eval_grid, None, grid + interp.output_size, derivative_orders, flags)


def eval_grid_(out, interp, *, derivative_orders=None, flags=0):
    """eval_grid!(out, interp; derivative_orders) — src/interpolation_parallel.jl:85-103."""
    grid = tuple(int(d.t_eval.shape[0]) for d in interp.interp_dims)
    return interp._eval(interp._h.lib.dind_eval_grid, out, grid + interp.output_size, derivative_orders, flags)
