// dind_fast_unstructured_aos.cu — eval_unstructured! for vector-valued u (2..4 output components)
// over a component-fastest ("AoS") copy of u.
//
// The reference stores u as (n_1..n_N, out...): every output component is a separate N-D volume,
// so a C-component stencil gather touches C distant sectors per node (config 5: 32 corners x 4
// components = 128 scattered 4-byte loads, ~6 KB of HBM traffic per iid point with 64-byte
// fetches).  The handle keeps a repacked copy u_aos[c + C*i] (one extra pass per dind_set_u, SURVEY.md
// §7.2 "repack allowed"): all components of a node are ONE 8/16/32-byte vector load (3-component nodes are
// padded to 4 so that a node never straddles a 32-byte sector).  Outputs keep the reference's layout (n_points, out...).
#include <cstdlib>

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int ABLOCK = 256;

template <typename T, int D, int W, int C, bool WIDE>
struct ContractV {
    static __device__ __forceinline__ Vec<T, C> run(const T* __restrict__ p, const T (&w)[MAXN][W], const int (&stride)[MAXN]) {
        Vec<T, C> acc;
#pragma unroll
        for (int c = 0; c < C; ++c) acc.v[c] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            const Vec<T, C> t = ContractV<T, D - 1, W, C, WIDE>::run(p + a * stride[D], w, stride);
#pragma unroll
            for (int c = 0; c < C; ++c) acc.v[c] = fma(w[D][a], t.v[c], acc.v[c]);
        }
        return acc;
    }
};
template <typename T, int W, int C, bool WIDE>
struct ContractV<T, 0, W, C, WIDE> {
    static __device__ __forceinline__ Vec<T, C> run(const T* __restrict__ p, const T (&w)[MAXN][W], const int (&)[MAXN]) {
        Vec<T, C> acc;
#pragma unroll
        for (int c = 0; c < C; ++c) acc.v[c] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            const Vec<T, C> t = load_node<T, C, WIDE>(p + a * (C == 3 ? 4 : C));
#pragma unroll
            for (int c = 0; c < C; ++c) acc.v[c] = fma(w[0][a], t.v[c], acc.v[c]);
        }
        return acc;
    }
};

// same per-dimension stencil as dind_fast_unstructured.cu
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
        else { w[0] = -r; w[1] = r; }
        return c;
    }
    const int idx = cached_idx ? cached_idx : clampi(dim_count<T, true>(D, x, kx), W, n - W);
    bspline_basis_ct<T, W - 1>(D.knots, D.recip, n, idx - 1, x, D.order, w);
    return idx - W;
}

// COH: the points come in plan order (cell-sorted).  Then every streaming input of the point is requested before
// the first one is used: the per-dimension stencils have control flow the compiler does not move loads across,
// and one t_eval load per dimension in sequence is N dependent DRAM round trips (ncu source view, config 5 in plan
// order: 34 % of all stall samples sat on them; 3.24 -> 2.54 ms).  With iid points the kernel is bound by the
// DRAM sector traffic of the gather and the early loads cost 5 % (config 5: 33.0 -> 34.9 ms), so that
// instantiation keeps the loads where they are used.
template <typename T, int N, int W, int C, bool WIDE, bool COH, int MINB = 0>
__global__ void __launch_bounds__(ABLOCK, MINB) unstructured_aos_kernel(const __grid_constant__ EvalParams<T> P) {
    const int64_t k = (int64_t)blockIdx.x * ABLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    constexpr int S = C == 3 ? 4 : C;   // aos_node_stride(C)
    T w[MAXN][W];
    int stride[MAXN];
    int64_t base = 0;
    const uint32_t cell = P.cells ? __ldcs(P.cells + k) : 0u;
    T xs[N];
    int64_t ko = k;
    if constexpr (COH) {
#pragma unroll
        for (int d = 0; d < N; ++d) xs[d] = __ldcs(P.dim[d].t_eval + k);
        if (P.perm) ko = (int64_t)__ldcs(P.perm + k);
    }
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        const T x = COH ? xs[d] : __ldcs(D.t_eval + k);
        const int cached = P.cells ? cell_idx<T>(D, cell) : (P.use_cached_idx ? __ldcs(D.idx_eval + k) : 0);
        const int i0 = stencil_ct<T, W>(D, x, cached, w[d]);
        stride[d] = (int)D.u_stride * S;
        base += (int64_t)i0 * D.u_stride;
    }
    if constexpr (!COH) ko = P.perm ? (int64_t)__ldcs(P.perm + k) : k;
    const Vec<T, C> acc = ContractV<T, N - 1, W, C, WIDE>::run(P.u_aos + base * S, w, stride);
    if (P.staged_records) {   // cell-sorted plan: one aligned record per point
        store_record<T, C>(P.out + ko * P.out_point_stride, acc.v);
        return;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) __stcs(P.out + (int64_t)c * P.out_comp_stride + ko * P.out_point_stride, acc.v[c]);
}

template <typename T>
__global__ void __launch_bounds__(ABLOCK) aos_repack_kernel(const T* __restrict__ u, T* __restrict__ u_aos, int64_t n, int C) {
    const int64_t i = (int64_t)blockIdx.x * ABLOCK + threadIdx.x;
    if (i >= n) return;
    const int S = aos_node_stride(C);
    for (int c = 0; c < S; ++c) u_aos[i * S + c] = c < C ? __ldcs(u + i + (int64_t)c * n) : T(0);
}

template <typename T, int N, int W>
cudaError_t launch_aos(const EvalParams<T>& P, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_points + ABLOCK - 1) / ABLOCK);
    const bool narrow = sizeof(T) == 8 && P.coherent_points;   // only 32-byte nodes have two load shapes
    const bool coh = P.coherent_points != 0;
    switch (P.n_comp) {
        case 2:
            if (coh) unstructured_aos_kernel<T, N, W, 2, true, true><<<nb, ABLOCK, 0, s>>>(P);
            else unstructured_aos_kernel<T, N, W, 2, true, false><<<nb, ABLOCK, 0, s>>>(P);
            break;
        case 3:
            // config 3 in plan order: asking for 4 CTAs/SM (64 registers instead of the 60 ptxas picks on its own)
            // changes the schedule: 1.43 -> 1.33 ms; 3 CTAs (80 registers) 1.49, 5 (48) 1.36, 6 (40, spills) 1.58
            if (narrow && W == 4 && N == 3) unstructured_aos_kernel<T, N, W, 3, sizeof(T) != 8, true, 4><<<nb, ABLOCK, 0, s>>>(P);
            else if (coh) unstructured_aos_kernel<T, N, W, 3, sizeof(T) != 8, true><<<nb, ABLOCK, 0, s>>>(P);
            else unstructured_aos_kernel<T, N, W, 3, true, false><<<nb, ABLOCK, 0, s>>>(P);
            break;
        case 4:
            if (coh) unstructured_aos_kernel<T, N, W, 4, sizeof(T) != 8, true><<<nb, ABLOCK, 0, s>>>(P);
            else unstructured_aos_kernel<T, N, W, 4, true, false><<<nb, ABLOCK, 0, s>>>(P);
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

template <typename T>
cudaError_t launch_aos_repack(const T* u, T* u_aos, int64_t n, int n_comp, cudaStream_t s) {
    aos_repack_kernel<T><<<(unsigned)((n + ABLOCK - 1) / ABLOCK), ABLOCK, 0, s>>>(u, u_aos, n, n_comp);
    return cudaGetLastError();
}

template <typename T>
bool aos_supported(const EvalParams<T>& P) {
    const int N = P.n_in;
    if (P.wts || P.n_comp < 2 || P.n_comp > 4 || N < 1 || N > 6) return false;
    const int W = P.dim[0].width;
    for (int d = 0; d < N; ++d) {
        if (P.dim[d].width != W || P.dim[d].kind == CONSTANT) return false;
        if (P.dim[d].kind == BSPLINE && P.dim[d].degree != W - 1) return false;
    }
    if (P.u_comp_stride * aos_node_stride(P.n_comp) >= (1LL << 31)) return false;
    if (W == 2) return true;
    return (W == 3 || W == 4) && (N == 2 || N == 3);
}

template <typename T>
bool try_eval_fast_unstructured_aos(const EvalParams<T>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    if (!P.u_aos || !aos_supported(P)) return false;
    const int N = P.n_in, W = P.dim[0].width;
    *launches = 1;
#define DIND_AOS(NN, WW)                                       \
    if (N == NN && W == WW) {                                  \
        *err = launch_aos<T, NN, WW>(P, s);                    \
        *name = "unstructured_aos<N=" #NN ",W=" #WW ">";       \
        return true;                                           \
    }
    DIND_AOS(1, 2) DIND_AOS(2, 2) DIND_AOS(3, 2) DIND_AOS(4, 2) DIND_AOS(5, 2) DIND_AOS(6, 2)
    DIND_AOS(2, 3) DIND_AOS(3, 3) DIND_AOS(2, 4) DIND_AOS(3, 4)
#undef DIND_AOS
    return false;
}

#define DIND_INST(T)                                                                                              \
    template cudaError_t launch_aos_repack<T>(const T*, T*, int64_t, int, cudaStream_t);                         \
    template bool aos_supported<T>(const EvalParams<T>&);                                                         \
    template bool try_eval_fast_unstructured_aos<T>(const EvalParams<T>&, cudaStream_t, const char**, cudaError_t*, int*);
DIND_INST(float)
DIND_INST(double)

}  // namespace dind
