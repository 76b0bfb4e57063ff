// dind_api.cu — the C ABI of libdind.so (include/dind.h): handle lifecycle, validation with
// the reference's assert semantics, device-resident layout of knots / t_eval / u, kernel
// dispatch.  There is no CPU evaluation path in this library.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "dind_handle.h"

using namespace dind;

namespace dind {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

void comm_release(dind_interp_s* h);   // dind_comm.cu

}  // namespace dind

namespace {

// internal: `out` is pinned host memory mapped into the device address space (small-batch path)
constexpr unsigned FLAG_OUT_MAPPED = 1u << 31;
// internal: stop after the caches / tables / plan are built (dind_build_caches)
constexpr unsigned FLAG_BUILD_ONLY = 1u << 30;
constexpr unsigned DIND_FLAG_PUBLIC_MASK = ~(FLAG_OUT_MAPPED | FLAG_BUILD_ONLY);
constexpr size_t SMALL_BYTES = 64 << 10;    // staging size of the small-batch path
constexpr int64_t SMALL_POINTS = 2048;
constexpr size_t REC_MIN_TABLE_BYTES = 96u << 20;   // tables below this stay L2-resident: the plain gather is fine
constexpr size_t REC_MAX_BYTES = 32ull << 30;       // cap of the cell-record copy of u

Knobs read_knobs() {
    auto geti = [](const char* name, int dflt) {
        const char* v = getenv(name);
        return v ? atoi(v) : dflt;
    };
    Knobs k;
    k.chunk = geti("DIND_CHUNK", 0);
    k.pf = geti("DIND_PF", 1);
    k.pf1 = geti("DIND_PF1", -1);
    k.l2_warm = geti("DIND_L2_WARM", -1);
    k.c3_tile = geti("DIND_C3_TILE", -1);
    k.no_lut = getenv("DIND_NO_LUT") != nullptr;
    k.no_split = getenv("DIND_NO_SPLIT") != nullptr;
    k.no_cells = getenv("DIND_NO_CELLS") != nullptr;
    k.no_aos = getenv("DIND_NO_AOS") != nullptr;
    k.no_small = getenv("DIND_NO_SMALL") != nullptr;
    k.no_slice = getenv("DIND_NO_SLICE") != nullptr;
    k.persist = getenv("DIND_PERSIST") != nullptr;
    k.rec = geti("DIND_REC", -1);
    k.no_plan_recs = getenv("DIND_NO_PLAN_RECS") != nullptr;
    k.no_batch_vec = getenv("DIND_NO_BATCH_VEC") != nullptr;
    k.no_den_cache = getenv("DIND_NO_DEN_CACHE") != nullptr;
    k.plan_buckets = geti("DIND_PLAN_BUCKETS", -1);
    k.tile_smem = geti("DIND_TILE_SMEM", -1);
    return k;
}

// dev_out (optional): the ordinal of the device that owns a cudaMalloc'ed pointer, -1 for host / managed memory
bool is_device_ptr(const void* p, int* dev_out = nullptr) {
    cudaPointerAttributes a;
    if (dev_out) *dev_out = -1;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (a.type == cudaMemoryTypeDevice && dev_out) *dev_out = a.device;
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// A device pointer handed to a handle must live on the handle's device (the kernels dereference it there).
int check_ptr_device(const dind_interp_s* h, const void* p, const char* what);


template <typename T>
int upload_dim(dind_interp_s* h, DimState& D) {
    using K = typename KeyOf<T>::type;
    const int n = (int)D.knots.size();
    std::vector<T> kn(n);
    std::vector<K> keys(n);
    for (int i = 0; i < n; ++i) {
        kn[i] = (T)D.knots[i];
        keys[i] = to_key(kn[i]);
    }
    std::vector<T> recip;
    if (D.kind == LINEAR) {
        recip.resize(n > 1 ? n - 1 : 1);
        for (int c = 0; c + 1 < n; ++c) recip[c] = T(1) / (kn[c + 1] - kn[c]);
    } else if (D.kind == BSPLINE) {
        const int p = D.degree;
        recip.assign((size_t)(p > 0 ? p : 1) * n, T(0));
        for (int d = 1; d <= p; ++d)
            for (int i = 0; i + d < n; ++i) {
                const T span = kn[i + d] - kn[i];
                recip[(size_t)(d - 1) * n + i] = (span == T(0)) ? T(0) : T(1) / span;   // safe_div, spline_utils.jl:22
            }
    } else {
        recip.assign(1, T(0));
    }
    CUDA_TRY(D.d_knots.ensure(n * sizeof(T)));
    CUDA_TRY(D.d_keys.ensure(n * sizeof(K)));
    CUDA_TRY(D.d_recip.ensure(recip.size() * sizeof(T)));
    CUDA_TRY(cudaMemcpyAsync(D.d_knots.p, kn.data(), n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(D.d_keys.p, keys.data(), n * sizeof(K), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(D.d_recip.p, recip.data(), recip.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    // bucket table for the interval search: ~4 buckets per knot, built on the device with the
    // same arithmetic the kernels use (dind_device.cuh: lut_bucket)
    D.lut_n = 0;
    if (n >= 8 && !h->knobs.no_lut) {
        int nb = 32;
        while (nb < 4 * n && nb < 4096) nb <<= 1;
        const T t0 = kn[0];
        const T scale = T(nb) / (kn[n - 1] - kn[0]);
        if (std::isfinite((double)scale) && scale > T(0)) {
            CUDA_TRY(D.d_lut.ensure((size_t)nb * sizeof(int2)));
            CUDA_TRY(launch_lut<T>((const T*)D.d_knots.p, n, t0, scale, nb, (int2*)D.d_lut.p, h->stream));
            D.lut_n = nb;
            D.lut_t0 = (double)t0;
            D.lut_scale = (double)scale;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // host vectors go out of scope
    return DIND_OK;
}

template <typename T>
DimParams<T> make_dim_params(const dind_interp_s* h, const DimState& D, int dim, int order) {
    using K = typename KeyOf<T>::type;
    DimParams<T> P;
    memset(&P, 0, sizeof P);
    P.knots = (const T*)D.d_knots.p;
    P.keys = (const K*)D.d_keys.p;
    P.recip = (const T*)D.d_recip.p;
    P.t_eval = (const T*)D.d_teval;
    P.idx_eval = D.idx_valid ? (const int32_t*)D.d_idx.p : nullptr;
    P.tab_i0 = (const int32_t*)D.table.i0.p;
    P.tab_w = (const T*)D.table.w.p;
    P.tab_flag = (const uint8_t*)D.table.flag.p;
    P.lut = D.lut_n ? (const int2*)D.d_lut.p : nullptr;
    P.lut_t0 = (T)D.lut_t0;
    P.lut_scale = (T)D.lut_scale;
    P.lut_n = D.lut_n;
    P.n_eval = D.n_eval;
    int64_t stride = 1;
    for (int i = 0; i < dim; ++i) stride *= h->u_set ? h->u_dims[i] : 1;
    P.u_stride = stride;
    P.n_knots = (int32_t)D.knots.size();
    P.kind = D.kind;
    P.degree = D.degree;
    P.width = D.width;
    P.n_u = (int32_t)D.n_u;
    P.order = order;
    return P;
}

int check_handle(dind_handle_t h) {
    if (!h) return fail(DIND_ERR_INVALID_ARGUMENT, "null handle");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(DIND_ERR_CUDA, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
    return DIND_OK;
}

int check_ptr_device(const dind_interp_s* h, const void* p, const char* what) {
    int dev = -1;
    if (p && is_device_ptr(p, &dev) && dev >= 0 && dev != h->device)
        return fail(DIND_ERR_INVALID_ARGUMENT, "%s is memory of device %d, the handle lives on device %d", what, dev, h->device);
    return DIND_OK;
}

// One cudaPointerGetAttributes per pointer: is it device memory, and if so of the handle's device?
int classify_ptr(const dind_interp_s* h, const void* p, const char* what, bool* on_device) {
    int dev = -1;
    *on_device = p && is_device_ptr(p, &dev);
    if (*on_device && dev >= 0 && dev != h->device)
        return fail(DIND_ERR_INVALID_ARGUMENT, "%s is memory of device %d, the handle lives on device %d", what, dev, h->device);
    return DIND_OK;
}

int check_dim(dind_handle_t h, int dim) {
    if (dim < 0 || dim >= h->n_in) return fail(DIND_ERR_INVALID_ARGUMENT, "dimension %d out of range 0..%d", dim, h->n_in - 1);
    return DIND_OK;
}

template <typename T>
int ensure_idx(dind_interp_s* h, int dim) {
    DimState& D = h->dims[dim];
    if (D.idx_valid) return DIND_OK;
    CUDA_TRY(D.d_idx.ensure((size_t)D.n_eval * sizeof(int32_t)));
    DimParams<T> P = make_dim_params<T>(h, D, dim, 0);
    CUDA_TRY(launch_idx<T>(P, (int32_t*)D.d_idx.p, h->stream));
    h->launches += D.n_eval ? 1 : 0;
    D.idx_valid = true;
    return DIND_OK;
}

template <typename T>
int ensure_table(dind_interp_s* h, int dim, int order) {
    DimState& D = h->dims[dim];
    if (D.table.order == order) return DIND_OK;
    CUDA_TRY(D.table.i0.ensure((size_t)D.n_eval * sizeof(int32_t)));
    CUDA_TRY(D.table.w.ensure((size_t)D.n_eval * D.width * sizeof(T)));
    CUDA_TRY(D.table.flag.ensure((size_t)D.n_eval));
    DimParams<T> P = make_dim_params<T>(h, D, dim, order);
    CUDA_TRY(launch_axis_table<T>(P, order, (int32_t*)D.table.i0.p, (T*)D.table.w.p, (uint8_t*)D.table.flag.p, h->stream));
    h->launches += D.n_eval ? 1 : 0;
    D.table.order = order;
    h->den_epoch++;   // a rebuilt axis table invalidates the cached stage-1 result of the NURBS weights
    return DIND_OK;
}

// validate_size_u / validate_cache (src/interpolation_utils.jl:39-63)
int validate_interp(dind_interp_s* h) {
    for (int d = 0; d < h->n_in; ++d)
        if (!h->dims[d].set) return fail(DIND_ERR_STATE, "interpolation dimension %d has not been set (dind_set_dim)", d);
    if (!h->u_set) return fail(DIND_ERR_STATE, "u has not been set (dind_set_u)");
    for (int d = 0; d < h->n_in; ++d)
        if (h->u_dims[d] != h->dims[d].n_u)
            return fail(DIND_ERR_SIZE_MISMATCH,
                        h->dims[d].kind == BSPLINE
                            ? "Expected size(u, %d) == %lld based on the BSplineInterpolation properties, got %lld."
                            : "For the first N_in dimensions of u the length must match the t of the corresponding interpolation dimension (dim %d: expected %lld, got %lld).",
                        d + 1, (long long)h->dims[d].n_u, (long long)h->u_dims[d]);
    // NURBS weights with Constant dimensions mixed in are an extension (no reference semantics: the
    // weights multiply the tensor-product stencil weights); Linear dimensions are rejected — the
    // reference would log an @error and silently ignore the weights (interpolation_utils.jl:65-67).
    if (h->w_set)   // validate_cache (src/interpolation_utils.jl:56-63): size(weights) == size(u)[1:N_in], whatever the call order was
        for (int d = 0; d < h->n_in; ++d)
            if ((int)h->w_dims.size() != h->n_in || h->w_dims[d] != h->u_dims[d])
                return fail(DIND_ERR_SIZE_MISMATCH, "The size of the weights array must match the length of the first N_in dimensions of u.");
    if (h->w_set)
        for (int d = 0; d < h->n_in; ++d)
            if (h->dims[d].kind == LINEAR)
                return fail(DIND_ERR_INVALID_ARGUMENT, "NURBS weights need BSpline interpolation dimensions (dimension %d is Linear)", d + 1);
    return DIND_OK;
}

// validate_derivative_orders (src/interpolation_utils.jl:5-32)
int validate_orders(dind_interp_s* h, const int* orders_in, bool multi_point, int* orders) {
    for (int d = 0; d < h->n_in; ++d) {
        orders[d] = orders_in ? orders_in[d] : 0;
        if (orders[d] < 0) return fail(DIND_ERR_INVALID_ARGUMENT, "Derivative orders must me non-negative.");
    }
    if (multi_point)
        for (int d = 0; d < h->n_in; ++d)
            if (h->dims[d].kind == BSPLINE && orders[d] > h->dims[d].max_deriv)
                return fail(DIND_ERR_INVALID_ARGUMENT,
                            "For BSpline interpolation, when using multi-point evaluation the derivative orders cannot be larger "
                            "than the `max_derivative_order_eval` of the `BSplineInterpolationDimension` (dimension %d: %d > %d).",
                            d + 1, orders[d], h->dims[d].max_deriv);
    if (h->w_set)
        for (int d = 0; d < h->n_in; ++d)
            if (orders[d] != 0) return fail(DIND_ERR_INVALID_ARGUMENT, "Currently partial derivatives of NURBS are not supported.");
    return DIND_OK;
}

template <typename T>
int ensure_premultiplied(dind_interp_s* h) {
    if (!h->w_set || h->uw_valid) return DIND_OK;
    const int64_t n = h->comp_stride;
    CUDA_TRY(h->uw.ensure((size_t)n * h->n_comp * sizeof(T)));
    CUDA_TRY(launch_premultiply<T>((const T*)h->d_u, (const T*)h->d_w, (T*)h->uw.p, n, (int)h->n_comp, h->stream));
    h->launches += 1;
    h->uw_valid = true;
    return DIND_OK;
}

// Common evaluation driver.  t_ptrs: device pointers to the per-dimension evaluation points.
template <typename T>
int run_eval(dind_interp_s* h, void* out, const int* orders, bool grid, const void* const* t_dev,
             const int64_t* n_eval, bool have_idx_cache, unsigned flags, bool gradient = false) {
    const int N = h->n_in;
    int64_t n_points = 1;
    if (grid) for (int d = 0; d < N; ++d) n_points *= n_eval[d];
    else n_points = n_eval[0];
    if (n_points == 0) return DIND_OK;

    int rc = DIND_OK;
    EvalParams<T> P;
    memset(&P, 0, sizeof P);
    P.knobs = &h->knobs;
    bool const_deriv = false, zero_result = false;
    for (int d = 0; d < N; ++d) {
        DimState& D = h->dims[d];
        if (grid) {
            rc = ensure_table<T>(h, d, orders[d]);
            if (rc) return rc;
        }
        P.dim[d] = make_dim_params<T>(h, D, d, orders[d]);
        P.dim[d].t_eval = (const T*)t_dev[d];
        P.dim[d].n_eval = n_eval[d];
        P.dim[d].out_stride = 1;
        if (grid) for (int e = 0; e < d; ++e) P.dim[d].out_stride *= n_eval[e];
        if (!have_idx_cache) P.dim[d].idx_eval = nullptr;
        if (D.kind == CONSTANT && orders[d] > 0) const_deriv = true;
        if (D.kind == LINEAR && orders[d] > 1) zero_result = true;
    }
    const bool out_on_device = (flags & (FLAG_OUT_MAPPED | FLAG_BUILD_ONLY)) || is_device_ptr(out);
    // DIND_FLAG_ASYNC with a HOST `out` (pinned memory): the D2H copy is queued on the handle's copy
    // stream and the call returns; the caller reads `out` after dind_synchronize.  Two staging buffers
    // let the copy of step i overlap the H2D + kernel of step i + 1.
    const bool pipelined = (flags & DIND_FLAG_ASYNC) && !out_on_device;
    if (pipelined && !h->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_kernel, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
    const int sb = pipelined ? (h->out_toggle ^= 1) : 0;
    const size_t out_bytes = (size_t)n_points * h->n_comp * sizeof(T) * (gradient ? N + 1 : 1);
    T* d_out = (T*)out;
    const bool untouched_semantics = const_deriv && h->u_dims.size() > (size_t)N;
    if (!out_on_device) {
        if (pipelined) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_d2h[sb], 0));   // this buffer's previous D2H is done
        else if (h->copy_stream) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
        CUDA_TRY(h->out_stage[sb].ensure(out_bytes));
        d_out = (T*)h->out_stage[sb].p;
        // Constant-with-derivative leaves vector outputs untouched off the knots
        // (src/interpolation_methods.jl:49-55): keep the caller's values.
        if (untouched_semantics) CUDA_TRY(cudaMemcpyAsync(d_out, out, out_bytes, cudaMemcpyHostToDevice, h->stream));
    }
    P.wts = h->w_set ? (const T*)h->d_w : nullptr;
    P.n_in = N;
    P.u_comp_stride = h->comp_stride;
    P.n_comp = (int32_t)h->n_comp;
    // NURBS numerator u*w: a separate pass per dind_set_u, except where the kernel forms it while staging u
    bool split = grid && !(flags & DIND_FLAG_GENERIC) && !const_deriv && !zero_result && grid_split_supported<T>(P) &&
                 !h->knobs.no_split;
    {
        const void* before = h->split_scratch.p;
        if (split && h->split_scratch.ensure(grid_split_scratch_bytes<T>(P)) != cudaSuccess) {
            cudaGetLastError();   // no room for the stage-1 results (at most 2x the size of out): one-pass kernel instead
            split = false;
        }
        if (h->split_scratch.p != before) h->den_epoch++;   // new scratch: nothing cached in it
    }
    P.split_den_cached = split && h->w_set && h->split_den_epoch == h->den_epoch && !h->knobs.persist && !h->knobs.no_den_cache;
    P.u_is_raw = split && h->w_set && !h->uw_valid && grid_split_slice_ok<T>(P);
    if (!P.u_is_raw && (rc = ensure_premultiplied<T>(h))) return rc;
    P.u = (h->w_set && !P.u_is_raw) ? (const T*)h->uw.p : (const T*)h->d_u;
    P.out = d_out;
    P.n_points = n_points;
    P.u_comp_stride = h->comp_stride;
    P.out_comp_stride = h->out_comp_stride_override ? h->out_comp_stride_override : n_points;
    P.out_point_stride = 1;
    P.n_comp = (int32_t)h->n_comp;
    P.n_in = N;
    P.scalar_out = h->u_dims.size() == (size_t)N;
    P.const_deriv = const_deriv;
    P.zero_result = zero_result;
    P.use_cached_idx = (!grid && have_idx_cache && (flags & DIND_FLAG_CACHED_IDX)) ? 1 : 0;
    if (P.use_cached_idx) {
        for (int d = 0; d < N; ++d) {
            rc = ensure_idx<T>(h, d);
            if (rc) return rc;
            P.dim[d].idx_eval = (const int32_t*)h->dims[d].d_idx.p;
        }
    }

    int unperm_ct = 0, unperm_stride = 0;   // != 0: results are staged component-fastest and brought home after the kernel
    bool use_buckets = false;
    if (!grid && have_idx_cache && (flags & DIND_FLAG_CELL_SORTED)) {
        if (n_points >= (1LL << 32) || h->comp_stride >= (1LL << 32))
            return fail(DIND_ERR_UNSUPPORTED, "DIND_FLAG_CELL_SORTED needs fewer than 2^32 points and u elements per component");
        // packed cell words: every dimension's first node in its own bit field of one 32-bit word, when they fit
        // (Linear / BSpline dimensions; a Constant dimension's node is not its interval index)
        int cell_bits = 0;
        for (int d = 0; d < N && cell_bits >= 0; ++d) {
            if (P.dim[d].kind == CONSTANT || h->knobs.no_cells) { cell_bits = -1; break; }
            int b = 1;
            while ((1 << b) < P.dim[d].n_u) ++b;
            P.dim[d].cell_shift = cell_bits;
            P.dim[d].cell_mask = (1 << b) - 1;
            P.dim[d].cell_add = P.dim[d].kind == LINEAR ? 1 : P.dim[d].width;
            cell_bits += b;
            if (cell_bits > 32) cell_bits = -1;
        }
        if (!h->plan_valid || h->plan_n != n_points) {
            h->plan_has_cells = cell_bits > 0 && h->plan_cells.ensure((size_t)n_points * sizeof(uint32_t)) == cudaSuccess;
            if (!h->plan_has_cells) (void)cudaGetLastError();
            uint32_t* cells = h->plan_has_cells ? (uint32_t*)h->plan_cells.p : nullptr;
            CUDA_TRY(h->plan_perm.ensure((size_t)n_points * sizeof(uint32_t)));
            T* ts[MAXN];
            for (int d = 0; d < N; ++d) {
                CUDA_TRY(h->plan_t[d].ensure((size_t)n_points * sizeof(T)));
                ts[d] = (T*)h->plan_t[d].p;
            }
            size_t scratch_bytes = 0;
            CUDA_TRY(plan_sort_points<T>(P, nullptr, ts, cells, cell_bits, nullptr, &scratch_bytes, nullptr, h->stream));
            // sort scratch (keys in/out, values in, CUB temp) stays with the handle: a new t_eval of the same
            // size re-plans without cudaMalloc / cudaFree (which would also synchronise the device)
            CUDA_TRY(h->plan_scratch.ensure(scratch_bytes));
            // coordinate records for the one-read-per-point gather of the sorted t_eval: in the staging buffer of the
            // un-permutation (free at this point); without room for them the gathers run per dimension
            T* recs = nullptr;
            if (!h->knobs.no_plan_recs) {
                if (h->plan_stage.ensure((size_t)n_points * plan_rec_elems<T>(N) * sizeof(T)) == cudaSuccess) recs = (T*)h->plan_stage.p;
                else (void)cudaGetLastError();
            }
            CUDA_TRY(plan_sort_points<T>(P, (uint32_t*)h->plan_perm.p, ts, cells, cell_bits, h->plan_scratch.p, &scratch_bytes, recs, h->stream));
            h->launches += 2 + N;
            h->plan_valid = true;
            h->plan_items[0].valid = h->plan_items[1].valid = false;
            h->plan_buckets_valid = false;
            h->plan_n = n_points;
        }
        for (int d = 0; d < N; ++d) {
            P.dim[d].t_eval = (const T*)h->plan_t[d].p;
            P.dim[d].idx_eval = nullptr;
        }
        P.use_cached_idx = 0;
        P.coherent_points = 1;
        P.cells = (h->plan_has_cells && cell_bits > 0) ? (const uint32_t*)h->plan_cells.p : nullptr;
        P.perm = (flags & DIND_FLAG_PLAN_ORDER) ? nullptr : (const uint32_t*)h->plan_perm.p;
        // Results go back to their original positions through a staging buffer of one aligned record per point
        // (dind_plan.cu): the kernel writes the record, a second pass copies it home with coalesced reads and
        // writes.  A direct scatter costs a random read-modify-write of a 32-byte sector per point AND component; a
        // single gather pass over a plan-ordered staging buffer runs at DRAM-sector / TLB speed (measured: config 3
        // 0.81 ms vs 0.30 ms for 2e7 points, config 5 7 ms vs 1 ms for 1e8 points).
        // (Constant-with-derivative keeps the direct scatter: it must leave entries untouched.)
        if (P.perm && !const_deriv) {
            const int ct = (int)h->n_comp * (gradient ? N + 1 : 1);
            unperm_stride = plan_record_stride(ct, (int)sizeof(T));
            CUDA_TRY(h->plan_stage.ensure((size_t)n_points * unperm_stride * sizeof(T)));
            unperm_ct = ct;
            P.out_point_stride = unperm_stride;
            P.staged_records = 1;
            // Where the record lands.  Records of less than a 32-byte sector: slots ordered by (destination window,
            // plan position) — every window's region then fills front to back while the kernel sweeps the plan, so
            // the records that share a sector (and a line) meet in L2 before it is written back (config 5, 16-byte
            // records: 3.2 ms vs 5.4 ms with the records at their original index).  Whole-sector records go straight
            // to staged[original index]: no second sort, no slot / destination arrays, and the copy home is a plain
            // coalesced transpose (config 3: plan 18.8 -> 9.1 ms, fused gradient 13.1 -> 11.9 ms, values 4.52 -> 4.60 ms).
            use_buckets = h->knobs.plan_buckets >= 0 ? h->knobs.plan_buckets != 0 : unperm_stride * (int)sizeof(T) < 32;
            if (use_buckets) {
                if (!h->plan_buckets_valid) {
                    CUDA_TRY(h->plan_slot2.ensure((size_t)n_points * sizeof(uint32_t)));
                    CUDA_TRY(h->plan_dest2.ensure((size_t)n_points * sizeof(uint32_t)));
                    size_t sb2 = 0;
                    CUDA_TRY(plan_bucket_schedule(nullptr, nullptr, nullptr, n_points, nullptr, &sb2, h->stream));
                    CUDA_TRY(h->plan_scratch.ensure(sb2));
                    sb2 = h->plan_scratch.cap;
                    CUDA_TRY(plan_bucket_schedule((const uint32_t*)h->plan_perm.p, (uint32_t*)h->plan_slot2.p,
                                                  (uint32_t*)h->plan_dest2.p, n_points, h->plan_scratch.p, &sb2, h->stream));
                    h->launches += 1;
                    h->plan_buckets_valid = true;
                }
                P.perm = (const uint32_t*)h->plan_slot2.p;   // staging slot of plan position k
            }
            P.out = (T*)h->plan_stage.p;
            P.out_comp_stride = 1;
        }
    }

    const bool fast_ok = !(flags & DIND_FLAG_GENERIC) && !const_deriv && !zero_result;
    // vector-valued u: gather from the component-fastest copy (rebuilt once per dind_set_u)
    const bool use_aos = fast_ok && !grid && aos_supported<T>(P) && !h->knobs.no_aos;
    if (use_aos) {
        if (!h->u_aos_valid) {
            CUDA_TRY(h->u_aos.ensure((size_t)h->comp_stride * aos_node_stride((int)h->n_comp) * sizeof(T)));
            CUDA_TRY(launch_aos_repack<T>((const T*)h->d_u, (T*)h->u_aos.p, h->comp_stride, (int)h->n_comp, h->stream));
            h->launches += 1;
            h->u_aos_valid = true;
        }
        P.u_aos = (const T*)h->u_aos.p;
    }
    // multilinear table larger than L2 at iid points: gather from the cell-record copy (one contiguous block per
    // point instead of 2^(N-1) row segments).  Built when one evaluation already repays the pass that writes it
    // (the gather saves ~1 KB of DRAM traffic per point), kept until the next dind_set_u.
    if (use_aos && !gradient && h->knobs.rec != 0 && rec_supported<T>(P)) {
        if (!h->u_rec_valid && !h->u_rec_declined) {
            const size_t bytes = rec_table_bytes<T>(P);
            const size_t aos_bytes = (size_t)h->comp_stride * aos_node_stride((int)h->n_comp) * sizeof(T);
            const bool want = h->knobs.rec == 1 || (aos_bytes > REC_MIN_TABLE_BYTES && (double)n_points * 1024.0 >= (double)bytes);
            if (want) {
                size_t free_b = 0, total_b = 0;
                CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
                if (bytes <= REC_MAX_BYTES && bytes <= free_b / 2 && h->u_rec.ensure(bytes) == cudaSuccess) {
                    CUDA_TRY(launch_rec_build<T>(P, (T*)h->u_rec.p, h->stream));
                    h->launches += 1;
                    h->u_rec_valid = true;
                } else {
                    (void)cudaGetLastError();
                    h->u_rec_declined = true;
                }
            }
        }
        if (h->u_rec_valid) {
            P.u_rec = (const T*)h->u_rec.p;
            rec_fill_strides<T>(P);
        }
    }
    // cell-sorted plan: the work-item schedule of the tile kernels (<= pts consecutive points of one cell per thread)
    dind_interp_s::PlanItems* IT = nullptr;
    if (use_aos && P.cells) {
        const int pts = tile_points_per_item<T>(P, gradient);
        if (pts > 0) {
            IT = &h->plan_items[gradient ? 1 : 0];
            if (!IT->valid || IT->pts != pts) {
                size_t sb = 0;
                CUDA_TRY(plan_build_items(nullptr, n_points, pts, nullptr, nullptr, nullptr, &sb, h->stream));
                CUDA_TRY(h->plan_scratch.ensure(sb));
                sb = h->plan_scratch.cap;
                int64_t n_items = 0;
                CUDA_TRY(plan_build_items(P.cells, n_points, pts, nullptr, &n_items, h->plan_scratch.p, &sb, h->stream));
                CUDA_TRY(IT->items.ensure((size_t)(n_items + 1) * sizeof(uint32_t)));
                CUDA_TRY(plan_build_items(P.cells, n_points, pts, (uint32_t*)IT->items.p, &n_items, h->plan_scratch.p, &sb, h->stream));
                IT->n = n_items;
                IT->pts = pts;
                P.items = (const uint32_t*)IT->items.p;
                P.n_items = n_items;
                CUDA_TRY(IT->cells.ensure((size_t)n_items * sizeof(uint32_t)));
                T* it[MAXN];
                for (int d = 0; d < N; ++d) {
                    CUDA_TRY(IT->t[d].ensure((size_t)n_items * pts * sizeof(T)));
                    it[d] = (T*)IT->t[d].p;
                }
                CUDA_TRY(plan_item_layout<T>(P, pts, (uint32_t*)IT->cells.p, it, h->stream));
                IT->valid = true;
                h->launches += 5;
            }
            P.items = (const uint32_t*)IT->items.p;
            P.n_items = IT->n;
            P.item_pts = pts;
        }
    }
    if (flags & FLAG_BUILD_ONLY) {
        if (!(flags & DIND_FLAG_ASYNC)) CUDA_TRY(cudaStreamSynchronize(h->stream));
        return DIND_OK;
    }

    bool handled = false;
    cudaError_t err = cudaSuccess;
    const char* name = nullptr;
    int nlaunch = 0;
    if (gradient) {
        // one fused pass when the configuration is in the kernel's envelope, otherwise n_in + 1
        // ordinary evaluations into the n_in + 1 output slots
        if (fast_ok && IT) {
            EvalParams<T> Q = P;   // the tile kernels read the item-major copies of the plan's inputs
            Q.cells = (const uint32_t*)IT->cells.p;
            for (int d = 0; d < N; ++d) Q.dim[d].t_eval = (const T*)IT->t[d].p;
            handled = try_eval_fast_tile<T>(Q, true, h->stream, &name, &err, &nlaunch);
        }
        if (fast_ok && !handled) handled = try_eval_fast_gradient<T>(P, h->stream, &name, &err, &nlaunch);
        if (!handled) {
            const int64_t slot = P.out_comp_stride * h->n_comp;   // n_points * n_comp, or n_comp when staged in plan order
            for (int sidx = 0; sidx <= N && err == cudaSuccess; ++sidx) {
                EvalParams<T> Q = P;
                for (int d = 0; d < N; ++d) Q.dim[d].order = (d + 1 == sidx) ? 1 : 0;
                Q.out = P.out + sidx * slot;
                int nl = 0;
                bool ok = !(flags & DIND_FLAG_GENERIC) && try_eval_fast_unstructured<T>(Q, h->stream, &name, &err, &nl);
                if (!ok) { err = launch_eval_generic<T>(Q, false, h->stream); nl = 1; }
                nlaunch += nl;
            }
            name = "gradient_by_separate_evaluations";
            handled = true;
        }
    }
    if (use_aos && !handled && IT) {
        EvalParams<T> Q = P;   // the tile kernels read the item-major copies of the plan's inputs
        Q.cells = (const uint32_t*)IT->cells.p;
        for (int d = 0; d < N; ++d) Q.dim[d].t_eval = (const T*)IT->t[d].p;
        handled = try_eval_fast_tile<T>(Q, false, h->stream, &name, &err, &nlaunch);
    }
    if (use_aos && !handled && P.u_rec) handled = try_eval_fast_unstructured_rec<T>(P, h->stream, &name, &err, &nlaunch);
    if (use_aos && !handled) handled = try_eval_fast_unstructured_aos<T>(P, h->stream, &name, &err, &nlaunch);
    if (split && !handled) {
        err = launch_grid_split<T>(P, h->split_scratch.p, h->stream, &name, &nlaunch);
        handled = err != cudaErrorNotSupported;
        if (err == cudaSuccess && h->w_set) h->split_den_epoch = h->den_epoch;
        else h->split_den_epoch = 0;
        if (!handled) err = cudaSuccess;
    }
    if (fast_ok && !handled) {
        handled = grid ? try_eval_fast_grid<T>(P, h->stream, &name, &err, &nlaunch)
                       : try_eval_fast_unstructured<T>(P, h->stream, &name, &err, &nlaunch);
    }
    if (!handled) {
        err = launch_eval_generic<T>(P, grid, h->stream);
        name = grid ? "generic_grid" : "generic_unstructured";
        nlaunch = 1;
    }
    CUDA_TRY(err);
    h->last_kernel = name;
    h->launches += nlaunch;
    if (unperm_ct) {
        if (use_buckets)
            CUDA_TRY(plan_unbucket<T>((const T*)h->plan_stage.p, (const uint32_t*)h->plan_dest2.p, d_out, n_points, unperm_ct, unperm_stride, h->stream));
        else
            CUDA_TRY(plan_unstage<T>((const T*)h->plan_stage.p, d_out, n_points, unperm_ct, unperm_stride, h->stream));
        h->launches += 1;
    }
    if (pipelined) {
        CUDA_TRY(cudaEventRecord(h->ev_kernel, h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_kernel, 0));
        CUDA_TRY(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
        CUDA_TRY(cudaEventRecord(h->ev_d2h[sb], h->copy_stream));
    } else if (!out_on_device) {
        CUDA_TRY(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    } else if (!(flags & DIND_FLAG_ASYNC)) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return DIND_OK;
}

template <typename T>
int eval_cached(dind_interp_s* h, void* out, const int* orders_in, bool grid, unsigned flags) {
    int rc = validate_interp(h);
    if (rc) return rc;
    int orders[MAXN];
    rc = validate_orders(h, orders_in, true, orders);
    if (rc) return rc;
    const void* t_dev[MAXN];
    int64_t n_eval[MAXN];
    for (int d = 0; d < h->n_in; ++d) {
        if (!h->dims[d].teval_set) return fail(DIND_ERR_STATE, "t_eval of dimension %d has not been set (dind_set_t_eval)", d + 1);
        t_dev[d] = h->dims[d].d_teval;
        n_eval[d] = h->dims[d].n_eval;
    }
    if (!grid)
        for (int d = 1; d < h->n_in; ++d)
            if (n_eval[d] != n_eval[0])
                return fail(DIND_ERR_SIZE_MISMATCH, "The t_eval of all interpolation dimensions must have the same length as the first dimension of out.");
    if (!out && !(flags & FLAG_BUILD_ONLY)) return fail(DIND_ERR_INVALID_ARGUMENT, "out is NULL");
    if ((rc = check_ptr_device(h, out, "out"))) return rc;
    return run_eval<T>(h, out, orders, grid, t_dev, n_eval, true, flags);
}

template <typename T>
int eval_points(dind_interp_s* h, void* out, const void* const* t, int64_t n, const int* orders_in, unsigned flags) {
    int rc = validate_interp(h);
    if (rc) return rc;
    int orders[MAXN];
    rc = validate_orders(h, orders_in, false, orders);   // uncached call: no max_derivative_order_eval limit
    if (rc) return rc;
    if (n < 0) return fail(DIND_ERR_INVALID_ARGUMENT, "n must be non-negative");
    if (!out || !t) return fail(DIND_ERR_INVALID_ARGUMENT, "out / t is NULL");
    const void* t_dev[MAXN];
    int64_t n_eval[MAXN];
    // the latency path (a single point is ~14 us end to end): ONE attribute query per pointer
    bool out_dev = false, t_on_dev[MAXN];
    if ((rc = classify_ptr(h, out, "out", &out_dev))) return rc;
    bool all_host = !out_dev;
    for (int d = 0; d < h->n_in; ++d) {
        if (!t[d]) return fail(DIND_ERR_INVALID_ARGUMENT, "t[%d] is NULL", d);
        if ((rc = classify_ptr(h, t[d], "t", &t_on_dev[d]))) return rc;
        n_eval[d] = n;
        all_host = all_host && !t_on_dev[d];
    }
    flags &= DIND_FLAG_PUBLIC_MASK & ~DIND_FLAG_CACHED_IDX;
    // Small-batch (latency) path — the single-point callable interp(t...) and ODE right-hand sides
    // (SURVEY.md §8f-4): the coordinates and the results travel through one pinned, device-mapped
    // staging buffer, so a call is ONE kernel launch + one stream synchronisation and no copy call.
    const size_t t_bytes = (size_t)n * sizeof(T);
    const size_t t_slot = (t_bytes + 15) & ~size_t(15);
    const size_t out_bytes = (size_t)n * h->n_comp * sizeof(T);
    if (all_host && n > 0 && n <= SMALL_POINTS && !(flags & DIND_FLAG_ASYNC) && t_slot * h->n_in + out_bytes <= SMALL_BYTES &&
        !h->knobs.no_small) {   // experiment knob: force the copy path
        if (!h->small_host) CUDA_TRY(cudaHostAlloc(&h->small_host, SMALL_BYTES, cudaHostAllocMapped));
        char* base = (char*)h->small_host;
        for (int d = 0; d < h->n_in; ++d) {
            memcpy(base + d * t_slot, t[d], t_bytes);
            t_dev[d] = base + d * t_slot;
        }
        char* so = base + h->n_in * t_slot;
        memcpy(so, out, out_bytes);   // Constant-with-derivative leaves vector outputs untouched (interpolation_methods.jl:49-55)
        rc = run_eval<T>(h, so, orders, false, t_dev, n_eval, false, flags | DIND_FLAG_ASYNC | FLAG_OUT_MAPPED);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        memcpy(out, so, out_bytes);
        return DIND_OK;
    }
    for (int d = 0; d < h->n_in; ++d) {
        if (t_on_dev[d]) {
            t_dev[d] = t[d];
        } else {
            CUDA_TRY(h->pts_stage[d].ensure((size_t)n * sizeof(T)));
            CUDA_TRY(cudaMemcpyAsync(h->pts_stage[d].p, t[d], (size_t)n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
            t_dev[d] = h->pts_stage[d].p;
        }
    }
    return run_eval<T>(h, out, orders, false, t_dev, n_eval, false, flags);
}

template <typename T>
int eval_point_rows_t(dind_interp_s* h, void* out_rows, int64_t a, int64_t b, int64_t comp_stride, const int* orders_in) {
    int orders[MAXN];
    int rc = validate_orders(h, orders_in, true, orders);
    if (rc) return rc;
    const void* t_dev[MAXN];
    int64_t n_eval[MAXN];
    for (int d = 0; d < h->n_in; ++d) {
        t_dev[d] = (const char*)h->dims[d].d_teval + (size_t)a * sizeof(T);
        n_eval[d] = b - a;
    }
    h->out_comp_stride_override = comp_stride;
    rc = run_eval<T>(h, out_rows, orders, false, t_dev, n_eval, false, DIND_FLAG_ASYNC);
    h->out_comp_stride_override = 0;
    return rc;
}

}  // namespace

int dind::eval_point_rows(dind_interp_s* h, void* out_rows, int64_t a, int64_t b, int64_t comp_stride, const int* derivative_orders) {
    return h->dtype == DIND_F32 ? eval_point_rows_t<float>(h, out_rows, a, b, comp_stride, derivative_orders)
                                : eval_point_rows_t<double>(h, out_rows, a, b, comp_stride, derivative_orders);
}

// ============================================================================================
extern "C" {

const char* dind_last_error(void) { return g_last_error.c_str(); }
int dind_version(void) { return DIND_VERSION; }

int dind_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int dind_create(dind_handle_t* handle, int n_in, int dtype, int device) {
    if (!handle) return fail(DIND_ERR_INVALID_ARGUMENT, "handle pointer is NULL");
    *handle = nullptr;
    if (n_in < 1 || n_in > MAXN) return fail(DIND_ERR_UNSUPPORTED, "n_in must be in 1..%d, got %d", MAXN, n_in);
    if (dtype != DIND_F32 && dtype != DIND_F64) return fail(DIND_ERR_UNSUPPORTED, "dtype must be DIND_F32 or DIND_F64");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(DIND_ERR_NO_DEVICE, "no CUDA device available (%s); libdind has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(DIND_ERR_INVALID_ARGUMENT, "device %d out of range 0..%d", device, ndev - 1);
    CUDA_TRY(cudaSetDevice(device));
    dind_interp_s* h = new dind_interp_s();
    h->n_in = n_in;
    h->dtype = dtype;
    h->device = device;
    h->esz = dtype == DIND_F32 ? 4 : 8;
    h->knobs = read_knobs();
    *handle = h;
    return DIND_OK;
}

int dind_destroy(dind_handle_t h) {
    if (!h) return DIND_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int d = 0; d < MAXN; ++d) {
        h->dims[d].release();
        h->pts_stage[d].release();
    }
    h->u_owned.release(); h->w_owned.release(); h->uw.release(); h->u_aos.release(); h->u_rec.release(); h->split_scratch.release(); h->out_stage[0].release(); h->out_stage[1].release();
    if (h->small_host) cudaFreeHost(h->small_host);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_kernel) cudaEventDestroy(h->ev_kernel);
    for (int i = 0; i < 2; ++i) if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
    comm_release(h);
    h->plan_perm.release(); h->plan_stage.release(); h->plan_scratch.release(); h->plan_slot2.release(); h->plan_dest2.release(); h->plan_cells.release(); h->plan_items[0].release(); h->plan_items[1].release();
    for (int d = 0; d < MAXN; ++d) h->plan_t[d].release();
    delete h;
    return DIND_OK;
}

int dind_set_stream(dind_handle_t h, void* cuda_stream) {
    int rc = check_handle(h);
    if (rc) return rc;
    h->stream = (cudaStream_t)cuda_stream;
    return DIND_OK;
}

int dind_synchronize(dind_handle_t h) {
    int rc = check_handle(h);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->copy_stream) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    return DIND_OK;
}

int dind_set_dim(dind_handle_t h, int dim, int kind, const void* t, int64_t n_t, int degree,
                 const int64_t* multiplicities, int left) {
    int rc = check_handle(h);
    if (rc) return rc;
    if ((rc = check_dim(h, dim))) return rc;
    if (kind != DIND_LINEAR && kind != DIND_CONSTANT && kind != DIND_BSPLINE)
        return fail(DIND_ERR_INVALID_ARGUMENT, "unknown interpolation dimension kind %d", kind);
    if (!t || n_t < 1) return fail(DIND_ERR_INVALID_ARGUMENT, "t must hold at least one knot");
    if (n_t > (1 << 24)) return fail(DIND_ERR_UNSUPPORTED, "more than 2^24 knots per dimension are not supported");
    std::vector<double> tv((size_t)n_t);
    std::vector<char> tmp;
    const void* src = t;
    if ((rc = check_ptr_device(h, t, "t"))) return rc;
    if (is_device_ptr(t)) {
        tmp.resize((size_t)n_t * h->esz);
        CUDA_TRY(cudaMemcpy(tmp.data(), t, tmp.size(), cudaMemcpyDeviceToHost));
        src = tmp.data();
    }
    for (int64_t i = 0; i < n_t; ++i) tv[i] = h->dtype == DIND_F32 ? (double)((const float*)src)[i] : ((const double*)src)[i];
    // validate_t (src/interpolation_utils.jl:34-37): all(>(0), diff(t)) — NaN fails too
    for (int64_t i = 1; i < n_t; ++i)
        if (!(tv[i] - tv[i - 1] > 0)) return fail(DIND_ERR_INVALID_ARGUMENT, "The elements of t must be sorted and unique.");
    if (std::isnan(tv[0])) return fail(DIND_ERR_INVALID_ARGUMENT, "The elements of t must be sorted and unique.");

    // Validate and derive everything in locals; the dimension's state is replaced only when all of it succeeded,
    // so a failing call leaves the previous (consistent) dimension in place.
    int n_degree = 0, n_width = 0;
    int64_t n_nu = 0;
    std::vector<int64_t> mult;
    std::vector<double> knots;
    if (kind == DIND_BSPLINE) {
        if (degree < 0) return fail(DIND_ERR_INVALID_ARGUMENT, "degree must be non-negative");
        if (degree > DIND_MAX_DEGREE) return fail(DIND_ERR_UNSUPPORTED, "B-spline degree %d > %d is not supported", degree, DIND_MAX_DEGREE);
        n_degree = degree;
        if (multiplicities) mult.assign(multiplicities, multiplicities + n_t);
        else {   // open/clamped knot vector (src/interpolation_dimensions.jl:168-173)
            mult.assign((size_t)n_t, 1);
            mult[0] = degree + 1;
            mult[(size_t)n_t - 1] = degree + 1;
        }
        for (int64_t i = 0; i < n_t; ++i)   // :175
            if (mult[i] < 1 || mult[i] > degree + 1)
                return fail(DIND_ERR_INVALID_ARGUMENT, "All multiplicities must be between 1 and degree + 1.");
        for (int64_t i = 0; i < n_t; ++i)   // expand_knot_vector_kernel (src/spline_utils.jl:1-20)
            for (int64_t m = 0; m < mult[i]; ++m) knots.push_back(tv[i]);
        n_width = degree + 1;
        n_nu = (int64_t)knots.size() - degree - 1;   // get_n_basis_functions (:223-225)
        if (n_nu < degree + 1)
            return fail(DIND_ERR_INVALID_ARGUMENT, "knot vector too short for degree %d (%lld basis functions)", degree, (long long)n_nu);
    } else {
        if (n_t < 2) return fail(DIND_ERR_INVALID_ARGUMENT, "Linear/Constant dimensions need at least 2 knots");
        knots = tv;
        n_width = kind == DIND_LINEAR ? 2 : 1;
        n_nu = n_t;
    }
    DimState& D = h->dims[dim];
    D.set = false;   // from here on the device tables are being rewritten: a CUDA failure leaves the dimension unset
    h->den_epoch++;
    D.invalidate_caches();
    h->plan_valid = false;
    D.kind = kind;
    D.left = left;
    D.degree = n_degree;
    D.t = tv;
    D.mult = mult;
    D.knots = knots;
    D.width = n_width;
    D.n_u = n_nu;
    rc = h->dtype == DIND_F32 ? upload_dim<float>(h, D) : upload_dim<double>(h, D);
    if (rc) return rc;
    D.set = true;
    return DIND_OK;
}

int dind_set_t_eval(dind_handle_t h, int dim, const void* t_eval, int64_t n_eval, int max_derivative_order_eval,
                    unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    if ((rc = check_dim(h, dim))) return rc;
    if (n_eval < 0) return fail(DIND_ERR_INVALID_ARGUMENT, "n_eval must be non-negative");
    if (n_eval > 0 && !t_eval) return fail(DIND_ERR_INVALID_ARGUMENT, "t_eval is NULL");
    if (max_derivative_order_eval < 0) return fail(DIND_ERR_INVALID_ARGUMENT, "max_derivative_order_eval must be non-negative");
    if ((rc = check_ptr_device(h, t_eval, "t_eval"))) return rc;
    DimState& D = h->dims[dim];
    const size_t bytes = (size_t)n_eval * h->esz;
    if (n_eval > 0 && is_device_ptr(t_eval) && !(flags & DIND_FLAG_COPY)) {
        D.d_teval = t_eval;   // borrowed
    } else {
        CUDA_TRY(D.teval_owned.ensure(bytes));
        if (n_eval > 0)
            CUDA_TRY(cudaMemcpyAsync(D.teval_owned.p, t_eval, bytes,
                                     is_device_ptr(t_eval) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        D.d_teval = D.teval_owned.p;
    }
    D.n_eval = n_eval;
    D.max_deriv = max_derivative_order_eval;
    D.teval_set = true;
    D.invalidate_caches();
    h->den_epoch++;
    h->plan_valid = false;
    return DIND_OK;
}

int dind_set_u(dind_handle_t h, const void* u, const int64_t* dims, int ndims, unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!u || !dims) return fail(DIND_ERR_INVALID_ARGUMENT, "u / dims is NULL");
    if (ndims < h->n_in)   // src/DataInterpolationsND.jl:44
        return fail(DIND_ERR_SIZE_MISMATCH, "The number of dimensions of u must be at least the number of interpolation dimensions.");
    int64_t comp_stride = 1, n_comp = 1;
    for (int i = 0; i < ndims; ++i) {
        if (dims[i] < 1) return fail(DIND_ERR_SIZE_MISMATCH, "size(u, %d) must be positive", i + 1);
        if (i < h->n_in) comp_stride *= dims[i]; else n_comp *= dims[i];
    }
    if (n_comp > (1 << 20)) return fail(DIND_ERR_UNSUPPORTED, "more than 2^20 output components are not supported");
    for (int d = 0; d < h->n_in; ++d)
        if (h->dims[d].set && dims[d] != h->dims[d].n_u)
            return fail(DIND_ERR_SIZE_MISMATCH, "size(u, %d) is %lld, the interpolation dimension expects %lld.", d + 1,
                        (long long)dims[d], (long long)h->dims[d].n_u);
    if (h->w_set)
        for (int d = 0; d < h->n_in; ++d)
            if ((int)h->w_dims.size() != h->n_in || h->w_dims[d] != dims[d])   // validate_cache (src/interpolation_utils.jl:56-63)
                return fail(DIND_ERR_SIZE_MISMATCH, "The size of the weights array must match the length of the first N_in dimensions of u.");
    if ((rc = check_ptr_device(h, u, "u"))) return rc;
    const size_t bytes = (size_t)comp_stride * n_comp * h->esz;
    if (is_device_ptr(u) && !(flags & DIND_FLAG_COPY)) {
        h->d_u = u;   // borrowed
    } else {
        CUDA_TRY(h->u_owned.ensure(bytes));
        CUDA_TRY(cudaMemcpyAsync(h->u_owned.p, u, bytes, is_device_ptr(u) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        // host source: wait unless the caller promises (DIND_FLAG_ASYNC) to keep the — pinned — buffer unchanged
        // until the next dind_synchronize
        if (!is_device_ptr(u) && !(flags & DIND_FLAG_ASYNC)) CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->d_u = h->u_owned.p;
    }
    if (!h->u_set || h->u_dims.size() != (size_t)ndims || !std::equal(dims, dims + ndims, h->u_dims.begin())) h->plan_valid = false;
    h->u_dims.assign(dims, dims + ndims);
    h->comp_stride = comp_stride;
    h->n_comp = n_comp;
    h->u_set = true;
    h->uw_valid = false;
    h->u_aos_valid = false;
    h->u_rec_valid = h->u_rec_declined = false;
    return DIND_OK;
}

int dind_set_weights(dind_handle_t h, const void* weights, const int64_t* dims, int ndims, unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    h->uw_valid = false;
    h->den_epoch++;
    if (!weights) {
        h->w_set = false;
        h->d_w = nullptr;
        h->w_dims.clear();
        return DIND_OK;
    }
    if ((rc = check_ptr_device(h, weights, "weights"))) return rc;
    if (!dims || ndims != h->n_in)
        return fail(DIND_ERR_SIZE_MISMATCH, "The size of the weights array must match the length of the first N_in dimensions of u.");
    int64_t n = 1;
    for (int d = 0; d < ndims; ++d) {
        if (h->u_set && dims[d] != h->u_dims[d])   // validate_cache (src/interpolation_utils.jl:56-63)
            return fail(DIND_ERR_SIZE_MISMATCH, "The size of the weights array must match the length of the first N_in dimensions of u.");
        if (h->dims[d].set && dims[d] != h->dims[d].n_u)
            return fail(DIND_ERR_SIZE_MISMATCH, "size(weights, %d) is %lld, the interpolation dimension expects %lld.", d + 1,
                        (long long)dims[d], (long long)h->dims[d].n_u);
        n *= dims[d];
    }
    const size_t bytes = (size_t)n * h->esz;
    if (is_device_ptr(weights) && !(flags & DIND_FLAG_COPY)) {
        h->d_w = weights;
    } else {
        CUDA_TRY(h->w_owned.ensure(bytes));
        CUDA_TRY(cudaMemcpyAsync(h->w_owned.p, weights, bytes, is_device_ptr(weights) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        if (!is_device_ptr(weights)) CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->d_w = h->w_owned.p;
    }
    h->w_dims.assign(dims, dims + ndims);
    h->w_set = true;
    return DIND_OK;
}

int dind_eval_unstructured(dind_handle_t h, void* out, const int* derivative_orders, unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    flags &= DIND_FLAG_PUBLIC_MASK;
    return h->dtype == DIND_F32 ? eval_cached<float>(h, out, derivative_orders, false, flags)
                                : eval_cached<double>(h, out, derivative_orders, false, flags);
}

int dind_eval_grid(dind_handle_t h, void* out, const int* derivative_orders, unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    flags &= DIND_FLAG_PUBLIC_MASK;
    return h->dtype == DIND_F32 ? eval_cached<float>(h, out, derivative_orders, true, flags)
                                : eval_cached<double>(h, out, derivative_orders, true, flags);
}

extern "C++" {
namespace {
template <typename T>
int eval_gradient(dind_interp_s* h, void* out, unsigned flags) {
    int rc = validate_interp(h);
    if (rc) return rc;
    if (h->w_set) return fail(DIND_ERR_UNSUPPORTED, "Currently partial derivatives of NURBS are not supported.");
    int orders[MAXN] = {0};
    const void* t_dev[MAXN];
    int64_t n_eval[MAXN];
    for (int d = 0; d < h->n_in; ++d) {
        const DimState& D = h->dims[d];
        if (D.kind == CONSTANT) return fail(DIND_ERR_UNSUPPORTED, "dind_eval_unstructured_gradient does not take Constant dimensions");
        if (D.kind == BSPLINE && D.max_deriv < 1)
            return fail(DIND_ERR_INVALID_ARGUMENT, "BSpline dimension %d needs max_derivative_order_eval >= 1 for gradients", d + 1);
        if (!D.teval_set) return fail(DIND_ERR_STATE, "t_eval of dimension %d has not been set (dind_set_t_eval)", d + 1);
        t_dev[d] = D.d_teval;
        n_eval[d] = D.n_eval;
        if (n_eval[d] != h->dims[0].n_eval)
            return fail(DIND_ERR_SIZE_MISMATCH, "The t_eval of all interpolation dimensions must have the same length as the first dimension of out.");
    }
    if (!out) return fail(DIND_ERR_INVALID_ARGUMENT, "out is NULL");
    if ((rc = check_ptr_device(h, out, "out"))) return rc;
    return run_eval<T>(h, out, orders, false, t_dev, n_eval, true, flags & ~DIND_FLAG_CACHED_IDX, true);
}
}  // namespace
}  // extern "C++"

int dind_eval_unstructured_gradient(dind_handle_t h, void* out, unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    flags &= DIND_FLAG_PUBLIC_MASK;
    return h->dtype == DIND_F32 ? eval_gradient<float>(h, out, flags) : eval_gradient<double>(h, out, flags);
}

int dind_build_caches(dind_handle_t h, int grid, const int* derivative_orders, unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    flags = (flags & DIND_FLAG_PUBLIC_MASK) | FLAG_BUILD_ONLY;
    return h->dtype == DIND_F32 ? eval_cached<float>(h, nullptr, derivative_orders, grid != 0, flags)
                                : eval_cached<double>(h, nullptr, derivative_orders, grid != 0, flags);
}

int dind_eval_points(dind_handle_t h, void* out, const void* const* t, int64_t n, const int* derivative_orders,
                     unsigned flags) {
    int rc = check_handle(h);
    if (rc) return rc;
    return h->dtype == DIND_F32 ? eval_points<float>(h, out, t, n, derivative_orders, flags)
                                : eval_points<double>(h, out, t, n, derivative_orders, flags);
}

int dind_get_idx_eval(dind_handle_t h, int dim, int64_t* idx_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if ((rc = check_dim(h, dim))) return rc;
    DimState& D = h->dims[dim];
    if (!D.set || !D.teval_set) return fail(DIND_ERR_STATE, "dimension %d needs dind_set_dim and dind_set_t_eval first", dim + 1);
    if (D.n_eval == 0) return DIND_OK;
    if (!idx_out) return fail(DIND_ERR_INVALID_ARGUMENT, "idx_out is NULL");
    rc = h->dtype == DIND_F32 ? ensure_idx<float>(h, dim) : ensure_idx<double>(h, dim);
    if (rc) return rc;
    std::vector<int32_t> tmp((size_t)D.n_eval);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), D.d_idx.p, tmp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < D.n_eval; ++i) idx_out[i] = tmp[i];
    return DIND_OK;
}

int dind_get_knots_all(dind_handle_t h, int dim, void* knots_out, int64_t* n_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if ((rc = check_dim(h, dim))) return rc;
    DimState& D = h->dims[dim];
    if (!D.set) return fail(DIND_ERR_STATE, "dimension %d has not been set", dim + 1);
    if (n_out) *n_out = (int64_t)D.knots.size();
    if (knots_out)
        for (size_t i = 0; i < D.knots.size(); ++i) {
            if (h->dtype == DIND_F32) ((float*)knots_out)[i] = (float)D.knots[i];
            else ((double*)knots_out)[i] = D.knots[i];
        }
    return DIND_OK;
}

int dind_get_basis_function_eval(dind_handle_t h, int dim, void* basis_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if ((rc = check_dim(h, dim))) return rc;
    DimState& D = h->dims[dim];
    if (!D.set || !D.teval_set) return fail(DIND_ERR_STATE, "dimension %d needs dind_set_dim and dind_set_t_eval first", dim + 1);
    if (D.kind != BSPLINE) return fail(DIND_ERR_INVALID_ARGUMENT, "dimension %d is not a BSpline dimension", dim + 1);
    if (D.n_eval == 0) return DIND_OK;
    if (!basis_out) return fail(DIND_ERR_INVALID_ARGUMENT, "basis_out is NULL");
    const size_t count = (size_t)D.n_eval * (D.degree + 1) * (D.max_deriv + 1);
    DevBuf tmp;
    CUDA_TRY(tmp.ensure(count * h->esz));
    cudaError_t e;
    if (h->dtype == DIND_F32) e = launch_basis_cache<float>(make_dim_params<float>(h, D, dim, 0), D.max_deriv, (float*)tmp.p, h->stream);
    else e = launch_basis_cache<double>(make_dim_params<double>(h, D, dim, 0), D.max_deriv, (double*)tmp.p, h->stream);
    h->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(basis_out, tmp.p, count * h->esz, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    tmp.release();
    CUDA_TRY(e);
    return DIND_OK;
}

int dind_get_plan_permutation(dind_handle_t h, uint32_t* perm_out) {
    int rc = check_handle(h);
    if (rc) return rc;
    if (!h->plan_valid) return fail(DIND_ERR_STATE, "no cell-sorted plan: evaluate with DIND_FLAG_CELL_SORTED first");
    if (!perm_out) return fail(DIND_ERR_INVALID_ARGUMENT, "perm_out is NULL");
    CUDA_TRY(cudaMemcpyAsync(perm_out, h->plan_perm.p, (size_t)h->plan_n * sizeof(uint32_t),
                             is_device_ptr(perm_out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return DIND_OK;
}

const char* dind_last_kernel(dind_handle_t h) { return h ? h->last_kernel.c_str() : "none"; }
int64_t dind_launch_count(dind_handle_t h) { return h ? h->launches : 0; }

int dind_shard_range(int64_t n, int rank, int world, int64_t* lo, int64_t* hi) {
    if (world < 1 || rank < 0 || rank >= world || n < 0 || !lo || !hi)
        return fail(DIND_ERR_INVALID_ARGUMENT, "bad shard arguments (n=%lld rank=%d world=%d)", (long long)n, rank, world);
    const int64_t q = n / world, r = n % world;
    *lo = rank * q + (rank < r ? rank : r);
    *hi = *lo + q + (rank < r ? 1 : 0);
    return DIND_OK;
}

int dind_host_register(void* ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(DIND_ERR_INVALID_ARGUMENT, "bad host_register arguments");
    CUDA_TRY(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return DIND_OK;
}

int dind_host_unregister(void* ptr) {
    if (!ptr) return fail(DIND_ERR_INVALID_ARGUMENT, "ptr is NULL");
    CUDA_TRY(cudaHostUnregister(ptr));
    return DIND_OK;
}

}  // extern "C"
