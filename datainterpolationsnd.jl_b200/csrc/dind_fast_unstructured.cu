// dind_fast_unstructured.cu — eval_unstructured! / batched interp(out, t...) with everything
// per-point in registers: branch-free interval search over the (L1-resident) knot keys,
// Cox–de Boor basis for a compile-time degree, then the W^N stencil gather contracted one
// axis at a time (axis 1 innermost = contiguous loads).
//
// Reference: eval_kernel -> _interpolate! (src/interpolation_parallel.jl:105-138,
// src/interpolation_methods.jl:1-141) reads idx_eval and basis_function_eval caches
// (32 B/point/dimension for a cubic) — here they are recomputed from the 4/8 B of t_eval,
// which is what makes the path HBM-lean.
//
// B200 mapping: one thread per point, 256-thread CTAs, grid sized to the point count.  The
// compulsory HBM traffic is t_eval in / out out; the u gather hits L1/L2 when u fits L2
// (configs 1 and 3) and HBM sectors otherwise (config 5).  FP32/FP64 FMA pipes; no tensor cores.
#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int UBLOCK = 256;

template <typename T, int D, int W>
struct ContractU {
    static __device__ __forceinline__ T run(const T* __restrict__ p, const T (&w)[MAXN][W], const int (&stride)[MAXN]) {
        T acc = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) acc = fma(w[D][a], ContractU<T, D - 1, W>::run(p + a * stride[D], w, stride), acc);
        return acc;
    }
};
template <typename T, int W>
struct ContractU<T, 0, W> {
    static __device__ __forceinline__ T run(const T* __restrict__ p, const T (&w)[MAXN][W], const int (&)[MAXN]) {
        T acc = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) acc = fma(w[0][a], __ldg(p + a), acc);
        return acc;
    }
};

// Stencil of one dimension whose width is the compile-time W.
template <typename T, int W>
__device__ __forceinline__ int stencil_ct(const DimParams<T>& D, T x, int cached_idx, T (&w)[W]) {
    using K = typename KeyOf<T>::type;
    const K kx = to_key(x);
    const int n = D.n_knots;
    if (W == 2 && D.kind == LINEAR) {
        const int idx = cached_idx ? cached_idx : clampi(dim_count<T, false>(D, x, kx), 1, n - 1);
        const int c = idx - 1;
        const T t1 = __ldg(D.knots + c), t2 = __ldg(D.knots + c + 1), r = __ldg(D.recip + c);
        if (D.order == 0) { w[0] = (t2 - x) * r; w[1] = (x - t1) * r; }
        else { w[0] = -r; w[1] = r; }   // order 1 (orders > 1 never reach this kernel)
        return c;
    }
    // BSpline of degree W-1
    const int idx = cached_idx ? cached_idx : clampi(dim_count<T, true>(D, x, kx), W, n - W);
    bspline_basis_ct<T, W - 1>(D.knots, D.recip, n, idx - 1, x, D.order, w);
    return idx - W;
}

// COH: plan order — all streaming inputs of the point are requested before the first one is used; the iid
// instantiation keeps each load at its use (see dind_fast_unstructured_aos.cu for the measurements)
template <typename T, int N, int W, bool NURBS, bool COH>
__global__ void __launch_bounds__(UBLOCK) unstructured_kernel(const __grid_constant__ EvalParams<T> P) {
    const int64_t k = (int64_t)blockIdx.x * UBLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    T w[MAXN][W];
    int stride[MAXN];
    int64_t base = 0;
    const uint32_t cell = P.cells ? __ldcs(P.cells + k) : 0u;
    T xs[N];
    int64_t ko = k;
    if constexpr (COH) {
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = __ldcs(P.dim[d].t_eval + k);
        if (P.perm) ko = (int64_t)__ldcs(P.perm + k);   // cell-sorted plan: output position
    }
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        const T x = COH ? xs[d] : __ldcs(D.t_eval + k);
        const int cached = P.cells ? cell_idx<T>(D, cell) : (P.use_cached_idx ? __ldcs(D.idx_eval + k) : 0);
        const int i0 = stencil_ct<T, W>(D, x, cached, w[d]);
        stride[d] = (int)D.u_stride;
        base += (int64_t)i0 * D.u_stride;
    }
    if constexpr (!COH) ko = P.perm ? (int64_t)__ldcs(P.perm + k) : k;
    T den = T(1);
    if (NURBS) den = ContractU<T, N - 1, W>::run(P.wts + base, w, stride);
    for (int c = 0; c < P.n_comp; ++c) {
        const T acc = ContractU<T, N - 1, W>::run(P.u + (int64_t)c * P.u_comp_stride + base, w, stride);
        __stcs(P.out + (int64_t)c * P.out_comp_stride + ko * P.out_point_stride, NURBS ? acc / den : acc);
    }
}

template <typename T, int N, int W>
cudaError_t launch_unstructured(const EvalParams<T>& P, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_points + UBLOCK - 1) / UBLOCK);
    if (P.coherent_points) {
        if (P.wts) unstructured_kernel<T, N, W, true, true><<<nb, UBLOCK, 0, s>>>(P);
        else unstructured_kernel<T, N, W, false, true><<<nb, UBLOCK, 0, s>>>(P);
    } else {
        if (P.wts) unstructured_kernel<T, N, W, true, false><<<nb, UBLOCK, 0, s>>>(P);
        else unstructured_kernel<T, N, W, false, false><<<nb, UBLOCK, 0, s>>>(P);
    }
    return cudaGetLastError();
}

}  // namespace

template <typename T>
bool try_eval_fast_unstructured(const EvalParams<T>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    const int N = P.n_in;
    if (N < 1 || N > 6) return false;
    const int W = P.dim[0].width;
    if (W < 2 || W > 6) return false;
    for (int d = 0; d < N; ++d) {
        if (P.dim[d].width != W) return false;
        if (P.dim[d].kind == CONSTANT) return false;
        if (P.dim[d].kind == BSPLINE && P.dim[d].degree != W - 1) return false;
    }
    if (P.u_comp_stride >= (1LL << 31)) return false;
    if ((P.n_points + UBLOCK - 1) / UBLOCK >= (1LL << 31)) return false;
    *launches = 1;
#define DIND_UNSTRUCTURED(NN, WW)                          \
    if (N == NN && W == WW) {                              \
        *err = launch_unstructured<T, NN, WW>(P, s);       \
        *name = "unstructured<N=" #NN ",W=" #WW ">";       \
        return true;                                       \
    }
    DIND_UNSTRUCTURED(1, 2) DIND_UNSTRUCTURED(1, 3) DIND_UNSTRUCTURED(1, 4) DIND_UNSTRUCTURED(1, 5) DIND_UNSTRUCTURED(1, 6)
    DIND_UNSTRUCTURED(2, 2) DIND_UNSTRUCTURED(2, 3) DIND_UNSTRUCTURED(2, 4) DIND_UNSTRUCTURED(2, 5) DIND_UNSTRUCTURED(2, 6)
    DIND_UNSTRUCTURED(3, 2) DIND_UNSTRUCTURED(3, 3) DIND_UNSTRUCTURED(3, 4)
    DIND_UNSTRUCTURED(4, 2) DIND_UNSTRUCTURED(4, 3)
    DIND_UNSTRUCTURED(5, 2)
    DIND_UNSTRUCTURED(6, 2)
#undef DIND_UNSTRUCTURED
    return false;
}

template bool try_eval_fast_unstructured<float>(const EvalParams<float>&, cudaStream_t, const char**, cudaError_t*, int*);
template bool try_eval_fast_unstructured<double>(const EvalParams<double>&, cudaStream_t, const char**, cudaError_t*, int*);

}  // namespace dind
