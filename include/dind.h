/* dind.h — C ABI of libdind.so: B200 (sm_100a) evaluation engine for the batched
 * evaluation hot path of DataInterpolationsND.jl.
 *
 * This header is the drop-in boundary.  The reference has no FFI: its seam
 * is Julia dispatch on the KernelAbstractions backend (`get_backend(out)` followed by
 * `synchronize(backend)`, src/interpolation_parallel.jl:40-50, :91-101) and
 * `Adapt.adapt_structure` as the host->device hook (src/DataInterpolationsND.jl:58-64).
 * A Julia host package `ccall`s the functions below instead (see INTEGRATION.md for the
 * binding); the Python package in datainterpolationsnd.jl_b200/ binds the same symbols with
 * ctypes.  Each entry point names the reference interface it replaces (paths relative to
 * the reference repository root).
 *
 * Conventions
 *  - Plain C types only; opaque handle; no exceptions cross the boundary.
 *  - Every function returns an int status: DIND_OK (0) or a negative DIND_ERR_* code;
 *    dind_last_error() returns a thread-local, human-readable message for the last failure
 *    on the calling thread.  The reference raises AssertionError (`@assert`) for every
 *    validation failure; the host wrapper turns a negative status into that exception.
 *  - Arrays are column-major (Julia): `u` is (n_1..n_N, out...), `out` is
 *    (n_points, out...) / (g_1..g_N, out...), element type = element type of `u`.
 *  - A data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer of the
 *    handle's device; the library detects which (cudaPointerGetAttributes).  Host data is
 *    copied; device data passed to dind_set_u / dind_set_t_eval / dind_set_weights is
 *    BORROWED (zero-copy, the analogue of `adapt_structure` handing over a device array):
 *    the caller keeps it alive and unchanged until it is replaced or the handle destroyed,
 *    unless DIND_FLAG_COPY is given.
 *  - All calls are synchronous like the reference's (each KA launch is followed by
 *    `synchronize(backend)`) unless DIND_FLAG_ASYNC is passed with a device `out`.
 *  - A handle may be used by one host thread at a time; distinct handles are independent.
 *  - There is no CPU fallback: without a CUDA device dind_create fails with
 *    DIND_ERR_NO_DEVICE.
 */
#ifndef DIND_H
#define DIND_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIND_VERSION 100 /* 0.1.0 */

/* status codes */
#define DIND_OK 0
#define DIND_ERR_INVALID_ARGUMENT (-1) /* a reference `@assert` on values would have failed      */
#define DIND_ERR_SIZE_MISMATCH (-2)    /* a reference `@assert` on sizes would have failed       */
#define DIND_ERR_UNSUPPORTED (-3)      /* valid in the reference, not supported by this engine   */
#define DIND_ERR_CUDA (-4)             /* CUDA runtime error, message has the CUDA error string  */
#define DIND_ERR_STATE (-5)            /* call order: e.g. eval before set_u / set_t_eval        */
#define DIND_ERR_NO_DEVICE (-6)        /* no usable CUDA device: there is no CPU fallback        */
#define DIND_ERR_NCCL (-7)             /* NCCL missing (libnccl.so.2 is bound at run time) or an NCCL call failed */

/* element types (eltype(u); knots, t_eval and out use the same type) */
#define DIND_F32 0
#define DIND_F64 1

/* interpolation dimension kinds: src/interpolation_dimensions.jl:15, :61, :119 */
#define DIND_LINEAR 0
#define DIND_CONSTANT 1
#define DIND_BSPLINE 2

/* flags */
#define DIND_FLAG_ASYNC 1u       /* do not wait before returning.  Device `out`: the kernel is queued on the stream.
                                   * Host `out` / host `u` (PINNED memory): the copies are queued too — D2H on the
                                   * handle's copy stream, overlapping the next step's H2D + kernel — and the
                                   * buffers belong to the library until dind_synchronize returns. */
#define DIND_FLAG_COPY 2u        /* copy device data instead of borrowing it                           */
#define DIND_FLAG_CACHED_IDX 4u  /* eval_unstructured: read the idx_eval cache instead of searching   */
#define DIND_FLAG_GENERIC 8u     /* force the generic (any-N, any-degree) kernels; testing/reference  */
#define DIND_FLAG_CELL_SORTED 16u /* eval_unstructured: evaluate through the cell-sorted plan — the point order
                                    * that makes the stencil gathers warp-coherent.  The plan (a permutation and
                                    * sorted copies of t_eval: n*(4 + n_in*sizeof(T)) bytes) is built on first use
                                    * and cached until t_eval, the knots or size(u) change — the analogue of the
                                    * reference's idx_eval / basis_function_eval caches.  Results are written to
                                    * their original positions: `out` is identical to the unsorted evaluation. */
#define DIND_FLAG_PLAN_ORDER 32u  /* with DIND_FLAG_CELL_SORTED: leave `out` in plan order (row k of out belongs to
                                    * original point perm[k], see dind_get_plan_permutation) instead of scattering
                                    * every result to its original position — the 4/8-byte scatter is what bounds
                                    * the sorted path (random read-modify-write of 32-byte sectors). */

#define DIND_MAX_DIMS 8   /* N_in <= 8 */
#define DIND_MAX_DEGREE 7 /* B-spline degree <= 7 */

typedef struct dind_interp_s* dind_handle_t;

/* ---- lifecycle --------------------------------------------------------------------------
 * Replaces: NDInterpolation(u, interp_dims; cache) — src/DataInterpolationsND.jl:29-56 — as
 * the owner of the device-resident state, and adapt_structure (:58-64) as the transfer hook.
 * `device` is a CUDA ordinal (one process per GPU: pass LOCAL_RANK). */
int dind_create(dind_handle_t* handle, int n_in, int dtype, int device);
int dind_destroy(dind_handle_t handle);
/* Use `cuda_stream` (a cudaStream_t; NULL = the legacy default stream) for every launch and copy. */
int dind_set_stream(dind_handle_t handle, void* cuda_stream);
int dind_synchronize(dind_handle_t handle);

/* ---- per-dimension setup ----------------------------------------------------------------
 * Replaces the dimension constructors LinearInterpolationDimension(t; t_eval),
 * ConstantInterpolationDimension(t; left, t_eval), BSplineInterpolationDimension(t, degree;
 * t_eval, max_derivative_order_eval, multiplicities) — src/interpolation_dimensions.jl:37-44,
 * :87-94, :162-204 — including validate_t (src/interpolation_utils.jl:34-37), the default
 * clamped multiplicities (:168-173), the multiplicity asserts (:174-175) and the knot
 * expansion kernel (src/spline_utils.jl:1-20).
 * `t`: n_t strictly increasing knots (host or device pointer; always copied).  A failing call leaves the
 * dimension as it was. `degree`, `multiplicities` (host, n_t
 * entries or NULL = clamped) only for DIND_BSPLINE. `left` is stored and, exactly like the
 * reference's field (src/interpolation_dimensions.jl:67), never read by evaluation. */
int dind_set_dim(dind_handle_t handle, int dim, int kind, const void* t, int64_t n_t, int degree,
                 const int64_t* multiplicities, int left);

/* Replaces the `t_eval` keyword of the constructors + set_eval_idx! / set_idx_kernel
 * (src/interpolation_utils.jl:143-161) + set_basis_function_eval!
 * (src/spline_utils.jl:164-191).  `t_eval`: host or device pointer, n_eval >= 0 points.
 * The per-point index cache is built lazily on the device (dind_get_idx_eval, or evaluation
 * with DIND_FLAG_CACHED_IDX); unstructured evaluation otherwise searches in-kernel and
 * recomputes the basis in registers, grid evaluation uses small per-axis tables. */
int dind_set_t_eval(dind_handle_t handle, int dim, const void* t_eval, int64_t n_eval,
                    int max_derivative_order_eval, unsigned flags);

/* ---- data -------------------------------------------------------------------------------
 * Replaces the `u` argument of NDInterpolation and validate_size_u
 * (src/interpolation_utils.jl:39-52).  dims[0..ndims-1] = size(u); ndims >= n_in.  Call again
 * for every new `u` (cheap for device pointers: borrowed, plus one pre-multiply pass when
 * NURBS weights are set). */
int dind_set_u(dind_handle_t handle, const void* u, const int64_t* dims, int ndims, unsigned flags);
/* Replaces NURBSWeights(weights) as `cache` + validate_cache (src/spline_utils.jl:232-234,
 * src/interpolation_utils.jl:54-63).  weights == NULL removes them (EmptyCache). */
int dind_set_weights(dind_handle_t handle, const void* weights, const int64_t* dims, int ndims,
                     unsigned flags);

/* ---- evaluation -------------------------------------------------------------------------
 * derivative_orders: n_in non-negative ints (NULL = all zero), validated as
 * validate_derivative_orders(...; multi_point = true) — src/interpolation_utils.jl:5-32.
 * `out`: host or device pointer with room for n_points * prod(out dims) elements. */

/* eval_unstructured!(out, interp; derivative_orders) — src/interpolation_parallel.jl:34-52. */
int dind_eval_unstructured(dind_handle_t handle, void* out, const int* derivative_orders,
                           unsigned flags);
/* eval_grid!(out, interp; derivative_orders) — src/interpolation_parallel.jl:85-103. */
int dind_eval_grid(dind_handle_t handle, void* out, const int* derivative_orders, unsigned flags);
/* Batched form of the uncached in-place call interp(out, t...; derivative_orders) —
 * src/DataInterpolationsND.jl:85-94: t[d] points to n values for dimension d (all host or all
 * device); out is (n, out...).  Does not touch the handle's t_eval.  Small all-host batches
 * (n <= 2048, <= 64 KB) take a latency path: coordinates and results travel through one pinned,
 * device-mapped staging buffer — one kernel launch + one stream synchronisation, no copy calls
 * (the single-point callable interp(t...), src/DataInterpolationsND.jl:73-83). */
int dind_eval_points(dind_handle_t handle, void* out, const void* const* t, int64_t n,
                     const int* derivative_orders, unsigned flags);

/* Beyond the reference: value and ALL first partial derivatives of the
 * cached unstructured points in one pass — equivalent to n_in + 1 calls of dind_eval_unstructured
 * with derivative_orders = (0..0), e_1, .., e_{n_in} (what the reference needs for a gradient,
 * src/interpolation_parallel.jl:34-52), sharing the interval search, the basis and the gather of u.
 * out: (n_points, out..., n_in + 1) column-major; slot 0 = value, slot d = d/dt_d.
 * Linear and BSpline dimensions (BSpline needs max_derivative_order_eval >= 1); Constant dimensions
 * and NURBS weights return DIND_ERR_UNSUPPORTED.  Honours DIND_FLAG_ASYNC / DIND_FLAG_CELL_SORTED. */
int dind_eval_unstructured_gradient(dind_handle_t handle, void* out, unsigned flags);

/* Cache construction as an explicit, timeable step — the work the reference's
 * constructors do in set_eval_idx! (src/interpolation_utils.jl:143-161) and
 * set_basis_function_eval! (src/spline_utils.jl:164-191), i.e. everything that depends on
 * t_eval / u but not on the evaluation itself, so that the first evaluation does not pay for it:
 *   grid != 0: the per-axis index/weight tables eval_grid uses for `derivative_orders`;
 *   grid == 0: the component-fastest copy of u (vector-valued u) and, for multilinear tables larger than L2 with enough
 *              iid points to repay it, its cell-record copy; the per-point index cache with
 *              DIND_FLAG_CACHED_IDX, the cell-sorted plan with DIND_FLAG_CELL_SORTED.
 * Needs dind_set_dim, dind_set_t_eval and dind_set_u.  Evaluations build whatever is missing
 * lazily, so calling this is optional.  Honours DIND_FLAG_ASYNC. */
int dind_build_caches(dind_handle_t handle, int grid, const int* derivative_orders, unsigned flags);

/* ---- introspection (parity tests; fields the reference exposes on its structs) ---------- */
/* idx_eval, 1-based Int64 exactly as the reference stores it (src/interpolation_dimensions.jl:38). */
int dind_get_idx_eval(dind_handle_t handle, int dim, int64_t* idx_out /* host, n_eval */);
/* knots_all of a BSpline dimension; n_out receives its length; knots_out may be NULL. */
int dind_get_knots_all(dind_handle_t handle, int dim, void* knots_out /* host */, int64_t* n_out);
/* basis_function_eval[n_eval, degree+1, max_derivative_order_eval+1] (column-major), computed
 * on the device on demand (src/spline_utils.jl:174-191). */
int dind_get_basis_function_eval(dind_handle_t handle, int dim, void* basis_out /* host */);
/* The cell-sorted plan's permutation (n_points entries, host or device pointer): plan position k
 * holds original point perm[k].  Valid after an evaluation with DIND_FLAG_CELL_SORTED. */
int dind_get_plan_permutation(dind_handle_t handle, uint32_t* perm_out);
/* Name of the kernel family the last evaluation on this handle dispatched to. */
const char* dind_last_kernel(dind_handle_t handle);
/* Number of kernel launches issued by this handle so far. */
int64_t dind_launch_count(dind_handle_t handle);

/* ---- multi-GPU helpers --------------------------------------------------------------------
 * Sharding is by contiguous index range (unstructured points) or by slab along the last
 * interpolation axis (grids); each rank creates its own handle over its slice of t_eval.
 * [lo, hi) = the part of 0..n-1 owned by `rank` of `world`; ranges differ by at most 1. */
int dind_shard_range(int64_t n, int rank, int world, int64_t* lo, int64_t* hi);

/* Optional all-gather of the outputs of a sharded evaluation (north star (5); the reference has no multi-device
 * path — src/interpolation_parallel.jl:34-52, :85-103 evaluate on ONE backend).  One process per GPU; the handle
 * owns the NCCL communicator.  NCCL (libnccl.so.2) is bound at run time by the first of these calls, so a
 * single-GPU host needs no NCCL; DIND_ERR_NCCL if it cannot be loaded.
 *   dind_comm_get_unique_id: rank 0 creates the 128-byte id and hands it to the other ranks by any means (MPI.jl
 *     bcast, a file, torch.distributed, ...);
 *   dind_comm_init: collective over the `world` handles, each on its own device; replaces a previous communicator;
 *   dind_comm_destroy: also done by dind_destroy. */
#define DIND_UNIQUE_ID_BYTES 128
int dind_comm_get_unique_id(void* id_out /* host, DIND_UNIQUE_ID_BYTES */);
int dind_comm_init(dind_handle_t handle, const void* unique_id, int rank, int world);
int dind_comm_destroy(dind_handle_t handle);
/* rank / world of the handle's communicator (0 / 1 without one) and the version of the bound NCCL; any may be NULL */
int dind_comm_info(dind_handle_t handle, int* rank, int* world, int* nccl_version);
/* Collective.  Every rank passes its local output and receives the global one, both DEVICE arrays of the handle's
 * element type, column-major:
 *     out_local  (unit, n_local, n_comp)      n_local = hi - lo of dind_shard_range(n_total, rank, world)
 *     out_global (unit, n_total, n_comp)
 * eval_unstructured shards: unit = 1, n_total = global number of points, n_comp = prod(out dims)
 *     (x (n_in + 1) for dind_eval_unstructured_gradient);
 * eval_grid slabs: unit = g_1 * .. * g_{N-1}, n_total = global length of the last grid axis.
 * Every component column travels straight from out_local into its final place in out_global (one ncclAllGather per
 * column inside ncclGroupStart/End; ncclBroadcast per rank and column when n_total % world != 0): no staging, no
 * padding, no re-packing.  Runs on the handle's stream after whatever was queued there (DIND_FLAG_ASYNC: returns
 * without waiting).  Without a communicator (world = 1) it is a device copy (or nothing when the pointers are equal). */
int dind_allgather_out(dind_handle_t handle, const void* out_local, void* out_global, int64_t n_total, int64_t unit,
                       int64_t n_comp, unsigned flags);

/* Collective: dind_eval_unstructured of this rank's shard, written straight into its final rows of the GLOBAL output,
 * plus the all-gather of the other ranks' rows — the exchange of chunk j runs (NCCL point-to-point calls of one group,
 * on the communicator's own stream) while the kernel of chunk j + 1 computes, so the gather costs what exceeds the
 * evaluation instead of adding to it.  Still "after the kernels, never inside them" (north star (5)): a chunk is
 * exchanged when the kernel that produced it has finished.
 *   out_global: DEVICE array (n_total, out...) column-major, the same on every rank when the call returns;
 *   the handle's t_eval (dind_set_t_eval) holds the rank's shard: n_eval = hi - lo of dind_shard_range(n_total, rank, world);
 *   n_chunks <= 0: chosen from n_total / world (the same value on every rank).
 * No local output array, no staging: a GPU holds the global output exactly once.  Uneven shards need nothing special
 * (every message carries its own count).  Without a communicator (world = 1) this is dind_eval_unstructured into out_global.
 * The reference has no multi-device path (src/interpolation_parallel.jl:34-52 evaluates on ONE backend). */
int dind_eval_unstructured_allgather(dind_handle_t handle, void* out_global, int64_t n_total, const int* derivative_orders,
                                     int n_chunks, unsigned flags);

/* ---- pinned host memory (for the host-buffer path) ---------------------------------------- */
int dind_host_register(void* ptr, int64_t bytes);
int dind_host_unregister(void* ptr);

/* ---- misc ---------------------------------------------------------------------------------- */
const char* dind_last_error(void);
int dind_version(void);
int dind_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DIND_H */
