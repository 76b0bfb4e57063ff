// dind_fast_gradient.cu — value and all first partial derivatives of eval_unstructured in ONE pass.
//
// The reference obtains a gradient with N_in + 1 calls eval_unstructured!(...; derivative_orders =
// e_d) (BASELINE config 3: "plus first derivatives" = four calls, 192 B/point of compulsory traffic
// and four stencil gathers).  Search, knot loads and the (degree+1)^N gather of u do not depend on
// the derivative order, so this kernel computes per dimension the order-0 and order-1 stencil
// weights (w, dw) and contracts u once into N_in + 1 accumulators:
//     level 0:  g0 = sum_a w_1[a] u[a]            g1 = sum_a dw_1[a] u[a]
//     level d:  g_j = sum_a w_d[a] g_j^(a) (j <= d),  g_{d+1} = sum_a dw_d[a] g_0^(a)
// Output: out[(point, component, slot)], slot 0 = value, slot d = d/dt_d.
#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int GBLOCK = 256;

template <typename T, int M>
struct GVec {
    T g[M];
};

template <typename T, int D, int W>
struct ContractG {
    static __device__ __forceinline__ GVec<T, D + 2> run(const T* __restrict__ p, const T (&w)[MAXN][W], const T (&dw)[MAXN][W],
                                                         const int (&stride)[MAXN]) {
        GVec<T, D + 2> r;
#pragma unroll
        for (int j = 0; j < D + 2; ++j) r.g[j] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            const GVec<T, D + 1> sub = ContractG<T, D - 1, W>::run(p + a * stride[D], w, dw, stride);
#pragma unroll
            for (int j = 0; j < D + 1; ++j) r.g[j] = fma(w[D][a], sub.g[j], r.g[j]);
            r.g[D + 1] = fma(dw[D][a], sub.g[0], r.g[D + 1]);
        }
        return r;
    }
};
template <typename T, int W>
struct ContractG<T, 0, W> {
    static __device__ __forceinline__ GVec<T, 2> run(const T* __restrict__ p, const T (&w)[MAXN][W], const T (&dw)[MAXN][W],
                                                     const int (&)[MAXN]) {
        GVec<T, 2> r;
        r.g[0] = T(0);
        r.g[1] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            const T v = __ldg(p + a);
            r.g[0] = fma(w[0][a], v, r.g[0]);
            r.g[1] = fma(dw[0][a], v, r.g[1]);
        }
        return r;
    }
};

// The same contraction over the component-fastest copy of u: every node is ONE vector load for all C
// components (instead of C gathers from C distant volumes); accumulator (slot j, component c) = g[j * C + c].
template <typename T, int D, int W, int C, bool WIDE>
struct ContractGV {
    static __device__ __forceinline__ GVec<T, (D + 2) * C> run(const T* __restrict__ p, const T (&w)[MAXN][W],
                                                               const T (&dw)[MAXN][W], const int (&stride)[MAXN]) {
        GVec<T, (D + 2) * C> r;
#pragma unroll
        for (int j = 0; j < (D + 2) * C; ++j) r.g[j] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            const GVec<T, (D + 1) * C> sub = ContractGV<T, D - 1, W, C, WIDE>::run(p + a * stride[D], w, dw, stride);
#pragma unroll
            for (int j = 0; j < (D + 1) * C; ++j) r.g[j] = fma(w[D][a], sub.g[j], r.g[j]);
#pragma unroll
            for (int c = 0; c < C; ++c) r.g[(D + 1) * C + c] = fma(dw[D][a], sub.g[c], r.g[(D + 1) * C + c]);
        }
        return r;
    }
};
template <typename T, int W, int C, bool WIDE>
struct ContractGV<T, 0, W, C, WIDE> {
    static __device__ __forceinline__ GVec<T, 2 * C> run(const T* __restrict__ p, const T (&w)[MAXN][W], const T (&dw)[MAXN][W],
                                                         const int (&)[MAXN]) {
        GVec<T, 2 * C> r;
#pragma unroll
        for (int j = 0; j < 2 * C; ++j) r.g[j] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            const Vec<T, C> v = load_node<T, C, WIDE>(p + a * (C == 3 ? 4 : C));
#pragma unroll
            for (int c = 0; c < C; ++c) {
                r.g[c] = fma(w[0][a], v.v[c], r.g[c]);
                r.g[C + c] = fma(dw[0][a], v.v[c], r.g[C + c]);
            }
        }
        return r;
    }
};

// order-0 and order-1 stencil weights of one dimension (compile-time width)
template <typename T, int W>
__device__ __forceinline__ int stencil_w_dw(const DimParams<T>& D, T x, int cached_idx, T (&w)[W], T (&dw)[W]) {
    using K = typename KeyOf<T>::type;
    const K kx = to_key(x);
    const int n = D.n_knots;
    if (W == 2 && D.kind == LINEAR) {
        const int c = (cached_idx ? cached_idx : clampi(dim_count<T, false>(D, x, kx), 1, n - 1)) - 1;
        const T t1 = __ldg(D.knots + c), t2 = __ldg(D.knots + c + 1), r = __ldg(D.recip + c);
        w[0] = (t2 - x) * r; w[1] = (x - t1) * r;
        dw[0] = -r; dw[1] = r;
        return c;
    }
    const int idx = cached_idx ? cached_idx : clampi(dim_count<T, true>(D, x, kx), W, n - W);
    bspline_basis_ct<T, W - 1>(D.knots, D.recip, n, idx - 1, x, 0, w);
    bspline_basis_ct<T, W - 1>(D.knots, D.recip, n, idx - 1, x, 1, dw);
    return idx - W;
}

template <typename T, int N, int W>
__global__ void __launch_bounds__(GBLOCK) unstructured_gradient_kernel(const __grid_constant__ EvalParams<T> P) {
    const int64_t k = (int64_t)blockIdx.x * GBLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    T w[MAXN][W], dw[MAXN][W];
    int stride[MAXN];
    int64_t base = 0;
    const uint32_t cell = P.cells ? __ldcs(P.cells + k) : 0u;
    T xs[N];   // all streaming inputs requested up front (see dind_fast_unstructured_aos.cu)
#pragma unroll
    for (int d = 0; d < N; ++d) xs[d] = __ldcs(P.dim[d].t_eval + k);
    const int64_t ko = P.perm ? (int64_t)__ldcs(P.perm + k) : k;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        const int i0 = stencil_w_dw<T, W>(D, xs[d], P.cells ? cell_idx<T>(D, cell) : 0, w[d], dw[d]);
        stride[d] = (int)D.u_stride;
        base += (int64_t)i0 * D.u_stride;
    }
    const int64_t slot = P.out_comp_stride * P.n_comp;   // elements per (value | derivative) block
    for (int c = 0; c < P.n_comp; ++c) {
        const GVec<T, N + 1> g = ContractG<T, N - 1, W>::run(P.u + (int64_t)c * P.u_comp_stride + base, w, dw, stride);
        T* o = P.out + (int64_t)c * P.out_comp_stride + ko * P.out_point_stride;
#pragma unroll
        for (int j = 0; j < N + 1; ++j) __stcs(o + j * slot, g.g[j]);
    }
}

template <typename T, int N, int W, int C, bool WIDE>
__global__ void __launch_bounds__(GBLOCK, 2) unstructured_gradient_aos_kernel(const __grid_constant__ EvalParams<T> P) {
    const int64_t k = (int64_t)blockIdx.x * GBLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    constexpr int S = C == 3 ? 4 : C;   // aos_node_stride(C)
    T w[MAXN][W], dw[MAXN][W];
    int stride[MAXN];
    int64_t base = 0;
    const uint32_t cell = P.cells ? __ldcs(P.cells + k) : 0u;
    T xs[N];   // all streaming inputs requested up front (see dind_fast_unstructured_aos.cu)
#pragma unroll
    for (int d = 0; d < N; ++d) xs[d] = __ldcs(P.dim[d].t_eval + k);
    const int64_t ko = P.perm ? (int64_t)__ldcs(P.perm + k) : k;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        const int i0 = stencil_w_dw<T, W>(D, xs[d], P.cells ? cell_idx<T>(D, cell) : 0, w[d], dw[d]);
        stride[d] = (int)D.u_stride * S;
        base += (int64_t)i0 * D.u_stride;
    }
    const int64_t slot = P.out_comp_stride * C;
    const GVec<T, (N + 1) * C> g = ContractGV<T, N - 1, W, C, WIDE>::run(P.u_aos + base * S, w, dw, stride);
    T* o = P.out + ko * P.out_point_stride;
    if (P.staged_records) {   // cell-sorted plan: the record is g in this order (component fastest, then slot)
        store_record<T, (N + 1) * C>(o, g.g);
        return;
    }
#pragma unroll
    for (int j = 0; j < N + 1; ++j)
#pragma unroll
        for (int c = 0; c < C; ++c) __stcs(o + (int64_t)c * P.out_comp_stride + j * slot, g.g[j * C + c]);
}

template <typename T, int N, int W>
cudaError_t launch_gradient(const EvalParams<T>& P, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_points + GBLOCK - 1) / GBLOCK);
    // vector-valued u with a component-fastest copy at hand (the register budget allows (N + 1) * C <= 16 accumulators)
    if (P.u_aos && P.n_comp >= 2 && P.n_comp <= 4 && (N + 1) * P.n_comp <= 16 && P.u_comp_stride * aos_node_stride(P.n_comp) < (1LL << 31)) {
        const bool narrow = sizeof(T) == 8 && P.coherent_points;
        switch (P.n_comp) {
            case 2: unstructured_gradient_aos_kernel<T, N, W, 2, true><<<nb, GBLOCK, 0, s>>>(P); break;
            case 3:
                if constexpr ((N + 1) * 3 <= 16) {
                    if (narrow) unstructured_gradient_aos_kernel<T, N, W, 3, sizeof(T) != 8><<<nb, GBLOCK, 0, s>>>(P);
                    else unstructured_gradient_aos_kernel<T, N, W, 3, true><<<nb, GBLOCK, 0, s>>>(P);
                }
                break;
            default:
                if constexpr ((N + 1) * 4 <= 16) {
                    if (narrow) unstructured_gradient_aos_kernel<T, N, W, 4, sizeof(T) != 8><<<nb, GBLOCK, 0, s>>>(P);
                    else unstructured_gradient_aos_kernel<T, N, W, 4, true><<<nb, GBLOCK, 0, s>>>(P);
                }
                break;
        }
        return cudaGetLastError();
    }
    unstructured_gradient_kernel<T, N, W><<<nb, GBLOCK, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace

template <typename T>
bool try_eval_fast_gradient(const EvalParams<T>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    const int N = P.n_in;
    if (P.wts || N < 1 || N > 5) return false;
    const int W = P.dim[0].width;
    for (int d = 0; d < N; ++d) {
        if (P.dim[d].width != W || P.dim[d].kind == CONSTANT) return false;
        if (P.dim[d].kind == BSPLINE && P.dim[d].degree != W - 1) return false;
    }
    if (P.u_comp_stride >= (1LL << 31)) return false;
    *launches = 1;
#define DIND_GRAD(NN, WW)                                        \
    if (N == NN && W == WW) {                                    \
        *err = launch_gradient<T, NN, WW>(P, s);                 \
        *name = "unstructured_gradient<N=" #NN ",W=" #WW ">";    \
        return true;                                             \
    }
    DIND_GRAD(1, 2) DIND_GRAD(2, 2) DIND_GRAD(3, 2) DIND_GRAD(4, 2) DIND_GRAD(5, 2)
    DIND_GRAD(1, 3) DIND_GRAD(2, 3) DIND_GRAD(3, 3)
    DIND_GRAD(1, 4) DIND_GRAD(2, 4) DIND_GRAD(3, 4)
#undef DIND_GRAD
    return false;
}

template bool try_eval_fast_gradient<float>(const EvalParams<float>&, cudaStream_t, const char**, cudaError_t*, int*);
template bool try_eval_fast_gradient<double>(const EvalParams<double>&, cudaStream_t, const char**, cudaError_t*, int*);

}  // namespace dind
