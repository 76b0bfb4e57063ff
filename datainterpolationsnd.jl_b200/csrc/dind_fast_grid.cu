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

// Vector type for V consecutive outputs along g_1 (16-byte streaming stores).
template <typename T, int V> struct VecStore;
template <> struct VecStore<float, 4> {
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct VecStore<double, 2> {
    static __device__ __forceinline__ void st(double* p, const double (&v)[2]) { __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1])); }
};
template <typename T> struct VecStore<T, 1> {
    static __device__ __forceinline__ void st(T* p, const T (&v)[1]) { __stcs(p, v[0]); }
};

constexpr int MAXCHUNK = 64;

// One thread: V consecutive outputs along g_1 (one 16-byte store per step) at a fixed
// (g_2..g_{N-1}), walking `chunk` coordinates of the last axis.  The last axis' table slice
// for the chunk is staged in shared memory (warp-uniform LDS instead of L1 traffic).
template <typename T, int N, int W, int V, bool NURBS>
__global__ void __launch_bounds__(PBLOCK) grid_pencil_kernel(const __grid_constant__ EvalParams<T> P, int RV, int chunk) {
    __shared__ int s_i0[MAXCHUNK];
    __shared__ T s_w[MAXCHUNK * W];
    const DimParams<T>& L = P.dim[N - 1];
    const int g_lo = blockIdx.y * chunk;
    const int g_n = min((int)L.n_eval - g_lo, chunk);
    for (int i = threadIdx.x; i < g_n; i += PBLOCK) s_i0[i] = __ldg(L.tab_i0 + g_lo + i);
    for (int i = threadIdx.x; i < g_n * W; i += PBLOCK) s_w[i] = __ldg(L.tab_w + (int64_t)g_lo * W + i);
    __syncthreads();
    const int rv = blockIdx.x * PBLOCK + threadIdx.x;   // index over (n_1 / V) x n_2 x ... x n_{N-1}
    if (rv >= RV) return;

    T w0[V][W];            // axis-1 weights of the V outputs
    int off0[V];           // axis-1 first nodes
    T w[MAXN][W];          // axes 2..N-1 (w[0] unused)
    int stride[MAXN];
    int base = 0;
    const int n1v = (int)P.dim[0].n_eval / V;
    int rem = rv / n1v;
    const int g1 = (rv - rem * n1v) * V;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        off0[v] = __ldg(P.dim[0].tab_i0 + g1 + v);
#pragma unroll
        for (int a = 0; a < W; ++a) w0[v][a] = __ldg(P.dim[0].tab_w + (g1 + v) * W + a);
    }
    stride[0] = 1;
#pragma unroll
    for (int d = 1; d < N - 1; ++d) {
        const DimParams<T>& D = P.dim[d];
        const int ne = (int)D.n_eval;
        const int q = rem / ne;
        const int g = rem - q * ne;
        rem = q;
        stride[d] = (int)D.u_stride;
        base += __ldg(D.tab_i0 + g) * (int)D.u_stride;
#pragma unroll
        for (int a = 0; a < W; ++a) w[d][a] = __ldg(D.tab_w + g * W + a);
    }
    const int sN = (int)L.u_stride;
    const int64_t R = (int64_t)RV * V;

    // plane value for output v at last-axis node `node`
    auto plane = [&](const T* __restrict__ src, int v, int node) -> T {
        const T* p = src + (base + off0[v] + node * sN);
        if (N == 2) {
            T acc = T(0);
#pragma unroll
            for (int a = 0; a < W; ++a) acc = fma(w0[v][a], __ldg(p + a), acc);
            return acc;
        } else {
            // axes 2..N-1 outer, axis 1 innermost with this output's weights
            T acc = T(0);
            int idx[MAXN];
#pragma unroll
            for (int d = 0; d < MAXN; ++d) idx[d] = 0;
            constexpr int OUTER = (N >= 3 ? 1 : 0) * (N == 3 ? W : N == 4 ? W * W : N == 5 ? W * W * W : 1);
#pragma unroll
            for (int s = 0; s < OUTER; ++s) {
                int o = 0;
                T wp = T(1);
                int t = s;
#pragma unroll
                for (int d = 1; d < N - 1; ++d) {
                    const int a = t % W;
                    t /= W;
                    o += a * stride[d];
                    wp = (d == 1) ? w[d][a] : wp * w[d][a];
                }
                T row = T(0);
#pragma unroll
                for (int a = 0; a < W; ++a) row = fma(w0[v][a], __ldg(p + o + a), row);
                acc = fma(wp, row, acc);
            }
            return acc;
        }
    };

    for (int c = 0; c < P.n_comp; ++c) {
        const T* __restrict__ uc = P.u + (int64_t)c * P.u_comp_stride;
        T* __restrict__ op = P.out + (int64_t)c * P.out_comp_stride + (int64_t)rv * V + (int64_t)g_lo * R;
        T A[V][W], B[V][W];
        int cur = -(1 << 30);
        for (int gi = 0; gi < g_n; ++gi, op += R) {
            const int i0 = s_i0[gi];
            const int shift = i0 - cur;
            if (shift != 0) {
                if (shift > 0 && shift < W) {
                    for (int s = 0; s < shift; ++s) {   // slide the window one node at a time
#pragma unroll
                        for (int v = 0; v < V; ++v) {
#pragma unroll
                            for (int a = 0; a + 1 < W; ++a) { A[v][a] = A[v][a + 1]; if (NURBS) B[v][a] = B[v][a + 1]; }
                            A[v][W - 1] = plane(uc, v, cur + s + W);
                            if (NURBS) B[v][W - 1] = plane(P.wts, v, cur + s + W);
                        }
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v)
#pragma unroll
                        for (int a = 0; a < W; ++a) {
                            A[v][a] = plane(uc, v, i0 + a);
                            if (NURBS) B[v][a] = plane(P.wts, v, i0 + a);
                        }
                }
                cur = i0;
            }
            T wl[W];
#pragma unroll
            for (int a = 0; a < W; ++a) wl[a] = s_w[gi * W + a];
            T res[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                T acc = T(0), den = T(0);
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    acc = fma(wl[a], A[v][a], acc);
                    if (NURBS) den = fma(wl[a], B[v][a], den);
                }
                res[v] = NURBS ? acc / den : acc;
            }
            VecStore<T, V>::st(op, res);
        }
    }
}

template <typename T, int N, int W, int V>
cudaError_t launch_pencil_v(const EvalParams<T>& P, cudaStream_t s) {
    int64_t R = 1;
    for (int d = 0; d < N - 1; ++d) R *= P.dim[d].n_eval;
    const int RV = (int)(R / V);
    const int n_last = (int)P.dim[N - 1].n_eval;
    const int bx = (RV + PBLOCK - 1) / PBLOCK;
    // enough CTAs for >= ~8 per SM while keeping pencils long enough to amortise the window start-up
    int chunk = MAXCHUNK;
    while (chunk > 8 && (int64_t)bx * ((n_last + chunk - 1) / chunk) < 148 * 8) chunk >>= 1;
    dim3 grid((unsigned)bx, (unsigned)((n_last + chunk - 1) / chunk));
    if (P.wts) grid_pencil_kernel<T, N, W, V, true><<<grid, PBLOCK, 0, s>>>(P, RV, chunk);
    else grid_pencil_kernel<T, N, W, V, false><<<grid, PBLOCK, 0, s>>>(P, RV, chunk);
    return cudaGetLastError();
}

template <typename T, int N, int W>
cudaError_t launch_pencil(const EvalParams<T>& P, cudaStream_t s) {
    constexpr int V = 16 / sizeof(T);
    bool vec = P.dim[0].n_eval % V == 0 && (reinterpret_cast<uintptr_t>(P.out) % 16) == 0 && P.out_comp_stride % V == 0;
    if (vec) return launch_pencil_v<T, N, W, V>(P, s);
    return launch_pencil_v<T, N, W, 1>(P, s);
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
    if (P.u_comp_stride * (P.wts ? 1 : 1) >= (1LL << 31)) return false;
    int64_t R = 1;
    for (int d = 0; d < N - 1; ++d) R *= P.dim[d].n_eval;
    if (R >= (1LL << 31) || P.dim[N - 1].n_eval / 8 >= 65535) return false;
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
