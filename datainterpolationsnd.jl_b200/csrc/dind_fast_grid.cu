// dind_fast_grid.cu — eval_grid! as a separable contraction ("pencil" kernels).
//
// Reference: eval_grid! / eval_kernel (src/interpolation_parallel.jl:85-138) redoes the full
// (degree+1)^N tensor-product stencil for every grid point.  Here one thread owns one point r
// of the first N-1 grid axes (consecutive threads = consecutive g_1, so stores coalesce) and
// walks a chunk of the LAST axis.  It keeps a sliding window of W "plane values"
//     A[a] = sum over the stencils of axes 1..N-1 of  w_1 ... w_{N-1} * u[.., i0_N + a]
// in registers; an output is the W-term contraction of that window with the last axis'
// weights.  When t_eval of the last axis is sorted, consecutive outputs share the window or
// slide it by one node, so each u value is loaded once per pencil instead of once per output:
// for 2x-upsampled trilinear (BASELINE config 2) that is 2 loads + 2.5 FMA per output
// instead of 8 + 8.  Unsorted t_eval stays correct (the window is simply rebuilt).
//
// B200 mapping: HBM-bound on the output write (4.5 B/point algorithmic for config 2); u is read
// through L1/L2 (67 MB < 126 MB L2) with read-only loads, outputs leave with streaming
// stores so they do not evict u from L2.  Per-axis index/weight tables are a few KB and stay
// in L1.  Band widths are 2..4: FP32/FP64 FMA pipes, not tensor cores.
#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int PBLOCK = 128;

template <typename T>
__device__ __forceinline__ void store_streaming(T* p, T v) { __stcs(p, v); }

// Contraction over axes 0..D of the stencil rooted at p.
template <typename T, int D, int W>
struct Contract {
    static __device__ __forceinline__ T run(const T* __restrict__ p, const T (&w)[MAXN][W], const int (&stride)[MAXN]) {
        T acc = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) acc = fma(w[D][a], Contract<T, D - 1, W>::run(p + a * stride[D], w, stride), acc);
        return acc;
    }
};
template <typename T, int W>
struct Contract<T, 0, W> {
    static __device__ __forceinline__ T run(const T* __restrict__ p, const T (&w)[MAXN][W], const int (&)[MAXN]) {
        T acc = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) acc = fma(w[0][a], __ldg(p + a), acc);
        return acc;
    }
};

template <typename T, int N, int W, bool NURBS>
__global__ void __launch_bounds__(PBLOCK) grid_pencil_kernel(const __grid_constant__ EvalParams<T> P, int64_t R, int chunk) {
    const int64_t r = (int64_t)blockIdx.x * PBLOCK + threadIdx.x;
    if (r >= R) return;
    T w[MAXN][W];
    int stride[MAXN];
    int64_t base = 0;
    {
        int64_t rem = r;
#pragma unroll
        for (int d = 0; d < N - 1; ++d) {
            const DimParams<T>& D = P.dim[d];
            const int g = (int)(rem % D.n_eval);
            rem /= D.n_eval;
            stride[d] = (int)D.u_stride;
            base += (int64_t)__ldg(D.tab_i0 + g) * D.u_stride;
#pragma unroll
            for (int a = 0; a < W; ++a) w[d][a] = __ldg(D.tab_w + g * W + a);
        }
    }
    const DimParams<T>& L = P.dim[N - 1];
    const int64_t sN = L.u_stride;
    const int g_lo = blockIdx.y * chunk;
    const int g_hi = min((int)L.n_eval, g_lo + chunk);
    for (int c = 0; c < P.n_comp; ++c) {
        const T* __restrict__ uc = P.u + (int64_t)c * P.u_comp_stride + base;
        const T* __restrict__ wc = NURBS ? P.wts + base : nullptr;
        T* __restrict__ oc = P.out + (int64_t)c * P.out_comp_stride + r;
        T A[W], B[W];
        int cur = -(1 << 30);
        for (int g = g_lo; g < g_hi; ++g) {
            const int i0 = __ldg(L.tab_i0 + g);
            const int shift = i0 - cur;
            if (shift != 0) {
                if (shift > 0 && shift < W) {
                    for (int s = 0; s < shift; ++s) {   // slide the window one node at a time
#pragma unroll
                        for (int a = 0; a + 1 < W; ++a) { A[a] = A[a + 1]; if (NURBS) B[a] = B[a + 1]; }
                        const int64_t off = (int64_t)(cur + s + W) * sN;
                        A[W - 1] = Contract<T, N - 2, W>::run(uc + off, w, stride);
                        if (NURBS) B[W - 1] = Contract<T, N - 2, W>::run(wc + off, w, stride);
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < W; ++a) {
                        const int64_t off = (int64_t)(i0 + a) * sN;
                        A[a] = Contract<T, N - 2, W>::run(uc + off, w, stride);
                        if (NURBS) B[a] = Contract<T, N - 2, W>::run(wc + off, w, stride);
                    }
                }
                cur = i0;
            }
            T acc = T(0), den = T(0);
#pragma unroll
            for (int a = 0; a < W; ++a) {
                const T wl = __ldg(L.tab_w + g * W + a);
                acc = fma(wl, A[a], acc);
                if (NURBS) den = fma(wl, B[a], den);
            }
            store_streaming(oc + (int64_t)g * R, NURBS ? acc / den : acc);
        }
    }
}

template <typename T, int N, int W>
cudaError_t launch_pencil(const EvalParams<T>& P, cudaStream_t s) {
    int64_t R = 1;
    for (int d = 0; d < N - 1; ++d) R *= P.dim[d].n_eval;
    const int n_last = (int)P.dim[N - 1].n_eval;
    const int64_t bx = (R + PBLOCK - 1) / PBLOCK;
    // enough CTAs for >= ~8 per SM while keeping pencils long enough to amortise the window start-up
    int chunk = 64;
    while (chunk > 8 && bx * ((n_last + chunk - 1) / chunk) < 148 * 8) chunk >>= 1;
    dim3 grid((unsigned)bx, (unsigned)((n_last + chunk - 1) / chunk));
    if (P.wts) grid_pencil_kernel<T, N, W, true><<<grid, PBLOCK, 0, s>>>(P, R, chunk);
    else grid_pencil_kernel<T, N, W, false><<<grid, PBLOCK, 0, s>>>(P, R, chunk);
    return cudaGetLastError();
}

}  // namespace

template <typename T>
bool try_eval_fast_grid(const EvalParams<T>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    const int N = P.n_in;
    if (N < 2 || N > 5) return false;
    const int W = P.dim[0].width;
    for (int d = 0; d < N; ++d) {
        if (P.dim[d].width != W) return false;
        if (P.dim[d].n_eval >= (1 << 30)) return false;
    }
    if (W < 2 || W > 4) return false;
    if (P.u_comp_stride >= (1LL << 31)) return false;
    int64_t R = 1;
    for (int d = 0; d < N - 1; ++d) R *= P.dim[d].n_eval;
    if ((R + PBLOCK - 1) / PBLOCK >= (1LL << 31) || P.dim[N - 1].n_eval / 8 >= 65535) return false;
    *launches = 1;
#define DIND_PENCIL(NN, WW)                              \
    if (N == NN && W == WW) {                            \
        *err = launch_pencil<T, NN, WW>(P, s);           \
        *name = "grid_pencil<N=" #NN ",W=" #WW ">";      \
        return true;                                     \
    }
    DIND_PENCIL(2, 2) DIND_PENCIL(2, 3) DIND_PENCIL(2, 4)
    DIND_PENCIL(3, 2) DIND_PENCIL(3, 3) DIND_PENCIL(3, 4)
    DIND_PENCIL(4, 2) DIND_PENCIL(4, 3) DIND_PENCIL(4, 4)
    DIND_PENCIL(5, 2)
#undef DIND_PENCIL
    return false;
}

template bool try_eval_fast_grid<float>(const EvalParams<float>&, cudaStream_t, const char**, cudaError_t*, int*);
template bool try_eval_fast_grid<double>(const EvalParams<double>&, cudaStream_t, const char**, cudaError_t*, int*);

}  // namespace dind
