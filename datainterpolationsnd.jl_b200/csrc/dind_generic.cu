// dind_generic.cu — cache-building kernels and the generic (any N_in, any kinds, any degree)
// evaluation kernels.  The generic kernels are the semantic baseline on the device: every
// specialised kernel is tested against them and against the CPU oracle.
//
// Reference: eval_kernel (src/interpolation_parallel.jl:105-138), one work item per output
// point; set_idx_kernel (src/interpolation_utils.jl:156-161); basis_function_eval_kernel
// (src/spline_utils.jl:174-191).
#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int BLOCK = 256;

inline unsigned blocks_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

template <typename T>
__global__ void __launch_bounds__(BLOCK) idx_kernel(DimParams<T> D, int32_t* __restrict__ idx_out) {
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= D.n_eval) return;
    T w[MAXW];
    int i0, idx1;
    bool on_knot;
    dim_stencil<T>(D, D.t_eval[i], 0, 0, i0, w, on_knot, idx1);
    idx_out[i] = idx1;
}

template <typename T>
__global__ void __launch_bounds__(BLOCK) basis_cache_kernel(DimParams<T> D, int max_deriv, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= D.n_eval) return;
    const int W = D.degree + 1;
    for (int r = 0; r <= max_deriv; ++r) {
        T w[MAXW];
        int i0, idx1;
        bool on_knot;
        dim_stencil<T>(D, D.t_eval[i], r, 0, i0, w, on_knot, idx1);
        for (int k = 0; k < W; ++k) out[i + D.n_eval * (k + (int64_t)W * r)] = w[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(BLOCK) axis_table_kernel(DimParams<T> D, int order, int32_t* __restrict__ tab_i0,
                                                           T* __restrict__ tab_w, uint8_t* __restrict__ tab_flag) {
    const int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (g >= D.n_eval) return;
    T w[MAXW];
    int i0, idx1;
    bool on_knot;
    const int cached = D.idx_eval ? D.idx_eval[g] : 0;
    dim_stencil<T>(D, D.t_eval[g], order, cached, i0, w, on_knot, idx1);
    tab_i0[g] = i0;
    tab_flag[g] = on_knot ? 1 : 0;
    for (int a = 0; a < D.width; ++a) tab_w[g * D.width + a] = w[a];
}

template <typename T>
__global__ void __launch_bounds__(BLOCK) lut_kernel(const T* __restrict__ knots, int n, T t0, T scale, int nb, int2* __restrict__ lut) {
    const int b = blockIdx.x * BLOCK + threadIdx.x;
    if (b >= nb) return;
    // first knot whose bucket is >= b, and >= b + 1 (buckets are non-decreasing along the sorted knots)
    int first[2];
    for (int q = 0; q < 2; ++q) {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (lut_bucket(knots[m], t0, scale, nb) < b + q) lo = m + 1; else hi = m;
        }
        first[q] = lo;
    }
    lut[b] = make_int2(first[0], first[1]);
}

template <typename T>
__global__ void __launch_bounds__(BLOCK) premultiply_kernel(const T* __restrict__ u, const T* __restrict__ w,
                                                            T* __restrict__ uw, int64_t n, int n_comp) {
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const T wi = w[i];
    for (int c = 0; c < n_comp; ++c) uw[i + (int64_t)c * n] = u[i + (int64_t)c * n] * wi;
}

// One thread per output point.  GRID: the point index decomposes into per-axis grid
// coordinates (dim 1 fastest) and the stencil comes from the per-axis tables; otherwise the
// stencil is computed in-thread from t_eval[k] (search + basis in registers).
template <typename T, bool GRID>
__global__ void __launch_bounds__(BLOCK) eval_generic_kernel(const __grid_constant__ EvalParams<T> P) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    const int N = P.n_in;
    const int64_t ko = ((!GRID && P.perm) ? (int64_t)P.perm[k] : k) * (GRID ? 1 : P.out_point_stride);   // output position (cell-sorted plan)
    int i0[MAXN];
    T w[MAXN][MAXW];
    bool any_on_knot = false;
    if (GRID) {
        int64_t rem = k;
        for (int d = 0; d < N; ++d) {
            const DimParams<T>& D = P.dim[d];
            const int64_t g = rem % D.n_eval;
            rem /= D.n_eval;
            i0[d] = D.tab_i0[g];
            for (int a = 0; a < D.width; ++a) w[d][a] = D.tab_w[g * D.width + a];
            any_on_knot |= (D.tab_flag[g] & 1) != 0;
        }
    } else {
        for (int d = 0; d < N; ++d) {
            const DimParams<T>& D = P.dim[d];
            const int cached = P.use_cached_idx ? D.idx_eval[k] : 0;
            bool on_knot;
            int idx1;
            dim_stencil<T>(D, D.t_eval[k], D.order, cached, i0[d], w[d], on_knot, idx1);
            any_on_knot |= on_knot;
        }
    }
    if (P.const_deriv) {
        // src/interpolation_methods.jl:49-55: NaN on a knot, otherwise `out` is returned
        // unchanged — zero for scalar output (fresh make_out), untouched memory for N_out > 0.
        if (any_on_knot) {
            const T nan = T(0) / T(0);
            for (int c = 0; c < P.n_comp; ++c) P.out[ko + c * P.out_comp_stride] = nan;
        } else if (P.scalar_out) {
            P.out[ko] = T(0);
        }
        return;
    }
    if (P.zero_result) {  // Linear with a derivative order > 1 (:10)
        for (int c = 0; c < P.n_comp; ++c) P.out[ko + c * P.out_comp_stride] = T(0);
        return;
    }
    int64_t base = 0;
    int total = 1;
    for (int d = 0; d < N; ++d) {
        base += (int64_t)i0[d] * P.dim[d].u_stride;
        total *= P.dim[d].width;
    }
    T den = T(0);
    constexpr int CB = 4;
    for (int c0 = 0; c0 < P.n_comp; c0 += CB) {
        T acc[CB];
        for (int j = 0; j < CB; ++j) acc[j] = T(0);
        int I[MAXN];
        for (int d = 0; d < N; ++d) I[d] = 0;
        for (int s = 0; s < total; ++s) {
            T wp = w[0][I[0]];
            int64_t off = base + I[0] * P.dim[0].u_stride;
            for (int d = 1; d < N; ++d) {
                wp *= w[d][I[d]];
                off += I[d] * P.dim[d].u_stride;
            }
            if (P.wts && c0 == 0) den += wp * P.wts[off];
            for (int j = 0; j < CB; ++j)
                if (c0 + j < P.n_comp) acc[j] += wp * P.u[off + (int64_t)(c0 + j) * P.u_comp_stride];
            for (int d = 0; d < N; ++d) {
                if (++I[d] < P.dim[d].width) break;
                I[d] = 0;
            }
        }
        for (int j = 0; j < CB; ++j)
            if (c0 + j < P.n_comp) P.out[ko + (int64_t)(c0 + j) * P.out_comp_stride] = P.wts ? acc[j] / den : acc[j];
    }
}

}  // namespace

template <typename T>
cudaError_t launch_idx(const DimParams<T>& D, int32_t* idx_out, cudaStream_t s) {
    if (D.n_eval == 0) return cudaSuccess;
    idx_kernel<T><<<blocks_for(D.n_eval, BLOCK), BLOCK, 0, s>>>(D, idx_out);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_basis_cache(const DimParams<T>& D, int max_deriv, T* out, cudaStream_t s) {
    if (D.n_eval == 0) return cudaSuccess;
    basis_cache_kernel<T><<<blocks_for(D.n_eval, BLOCK), BLOCK, 0, s>>>(D, max_deriv, out);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_axis_table(const DimParams<T>& D, int order, int32_t* tab_i0, T* tab_w, uint8_t* tab_flag,
                              cudaStream_t s) {
    if (D.n_eval == 0) return cudaSuccess;
    axis_table_kernel<T><<<blocks_for(D.n_eval, BLOCK), BLOCK, 0, s>>>(D, order, tab_i0, tab_w, tab_flag);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_lut(const T* knots, int n_knots, T t0, T scale, int n_buckets, int2* lut, cudaStream_t s) {
    lut_kernel<T><<<blocks_for(n_buckets, BLOCK), BLOCK, 0, s>>>(knots, n_knots, t0, scale, n_buckets, lut);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_premultiply(const T* u, const T* w, T* uw, int64_t n, int n_comp, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    premultiply_kernel<T><<<blocks_for(n, BLOCK), BLOCK, 0, s>>>(u, w, uw, n, n_comp);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_eval_generic(const EvalParams<T>& P, bool grid, cudaStream_t s) {
    if (P.n_points == 0) return cudaSuccess;
    const unsigned nb = blocks_for(P.n_points, BLOCK);
    if (grid) eval_generic_kernel<T, true><<<nb, BLOCK, 0, s>>>(P);
    else eval_generic_kernel<T, false><<<nb, BLOCK, 0, s>>>(P);
    return cudaGetLastError();
}

#define DIND_INSTANTIATE(T)                                                                                    \
    template cudaError_t launch_lut<T>(const T*, int, T, T, int, int2*, cudaStream_t);                         \
    template cudaError_t launch_idx<T>(const DimParams<T>&, int32_t*, cudaStream_t);                          \
    template cudaError_t launch_basis_cache<T>(const DimParams<T>&, int, T*, cudaStream_t);                   \
    template cudaError_t launch_axis_table<T>(const DimParams<T>&, int, int32_t*, T*, uint8_t*, cudaStream_t); \
    template cudaError_t launch_premultiply<T>(const T*, const T*, T*, int64_t, int, cudaStream_t);           \
    template cudaError_t launch_eval_generic<T>(const EvalParams<T>&, bool, cudaStream_t);
DIND_INSTANTIATE(float)
DIND_INSTANTIATE(double)

}  // namespace dind
