// dind_fast_grid.cuh — eval_grid! as a separable contraction with register reuse ("pencil" kernel).
//
// Reference: eval_grid! / eval_kernel (src/interpolation_parallel.jl:85-138) redoes the full
// (degree+1)^N tensor-product stencil for every grid point.  Here a thread owns a small
// register tile of outputs — VX points along g_1 (strided by 32 so every warp-level load and
// store is one contiguous 128-byte line) x J points along g_2 — at fixed middle axes, and walks
// a chunk of the LAST grid axis.  It keeps a sliding window of W "plane values" per output
//     A[a] = sum over the stencils of axes 1..N-1 of  w_1 ... w_{N-1} * u[.., i0_N + a]
// in registers; an output is the W-term contraction of that window with the last axis'
// weights.  When t_eval of the last axis is sorted, consecutive outputs share the window or
// slide it by one node, so u is loaded once per pencil instead of once per output, and the J
// outputs along g_2 share the rows they read when their stencils overlap (REL below).  For
// 2x-upsampled trilinear (BASELINE config 2) this is ~1.3 loads + ~4.5 FMA per output instead
// of 8 + 8.  Unsorted or sparse t_eval stays correct (windows are rebuilt, rows not shared).
//
// B200 mapping: HBM-bound on the output write (4.5 B/point algorithmic for config 2); u is read
// through L1/L2 (67 MB < 126 MB L2) with read-only loads; outputs leave with streaming
// (evict-first) stores so they do not push u out of L2.  The last axis' index/weight table
// slice of the chunk is staged in shared memory (warp-uniform LDS).  Band widths are 2..4:
// FP32/FP64 FMA pipes, not tensor cores (a per-axis band matrix is >98 % zeros).
#pragma once
#include <cmath>
#include <initializer_list>
#include <cstdlib>
#include <cstring>

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int PBLOCK = 128;
constexpr int PWARPS = PBLOCK / 32;
constexpr int MAXCHUNK = 128;

inline const Knobs& knobs_of(const Knobs* k) {
    static const Knobs defaults;
    return k ? *k : defaults;
}

// Bulk L2 prefetch through the TMA unit (SASS: UBLKPF.L2); addr and bytes are multiples of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void* addr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(addr), "r"(bytes) : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* addr) {   // SASS: CCTL.E.PF1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(addr));
}

template <int W, int N>
struct MidCount { static constexpr int value = (N <= 3) ? 1 : (N == 4) ? W : W * W; };

// Everything a thread needs that does not depend on the last axis.
template <typename T, int N, int W, int VX, int J>
struct PencilCtx {
    static constexpr int MID = MidCount<W, N>::value;
    T w1[VX][W];       // axis-1 weights of the VX outputs
    T w2[J][W];        // axis-2 weights of the J outputs (N >= 3)
    T wm[MID];         // product of the middle axes' weights per middle stencil point
    int om[MID];       // ... and its offset in u
    int off1[VX];      // axis-1 first node of each output (+ offsets of the fixed axes)
    int row0;          // offset of output j=0's first row (axis 2)
    int rel_rows;      // first row of output j=1 relative to output j=0's
    int s1, sN;
    int pf1;           // L1 prefetch distance in nodes along the last axis (0 = off)
    int pf_off;        // this lane's share of the warp's plane footprint (one 128-byte line per lane), -1 = none
};

// REL: 0/1 = the J=2 outputs' axis-2 stencils start REL rows apart and share rows; 2 = unrelated rows.
template <int N, int W, int J, int REL>
struct RowCount { static constexpr int value = (N < 3) ? 1 : (J == 1) ? W : (REL == 2 ? 2 * W : W + REL); };

// Raw loads of one plane: MID x rows x VX outputs x W nodes along axis 1 ...
template <typename T, int N, int W, int VX, int J, int REL>
__device__ __forceinline__ void plane_load(const PencilCtx<T, N, W, VX, J>& C, const T* const (&pv)[VX], int node,
                                           T (&raw)[MidCount<W, N>::value][RowCount<N, W, J, REL>::value][VX][W]) {
    constexpr int MID = MidCount<W, N>::value;
    constexpr int NR = RowCount<N, W, J, REL>::value;
#pragma unroll
    for (int m = 0; m < MID; ++m)
#pragma unroll
        for (int row = 0; row < NR; ++row) {
            int ro = row * C.s1;
            if (N >= 3 && J == 2 && REL == 2 && row >= W) ro = (C.rel_rows + row - W) * C.s1;
            const int o = node * C.sN + C.om[m] + C.row0 + ro;
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                const T* p = pv[v] + o;
#pragma unroll
                for (int a = 0; a < W; ++a) raw[m][row][v][a] = __ldg(p + a);
            }
        }
}

// ... and their reduction to the plane values of the J x VX outputs.
template <typename T, int N, int W, int VX, int J, int REL>
__device__ __forceinline__ void plane_reduce(const PencilCtx<T, N, W, VX, J>& C,
                                             const T (&raw)[MidCount<W, N>::value][RowCount<N, W, J, REL>::value][VX][W],
                                             T (&res)[J][VX]) {
    constexpr int MID = MidCount<W, N>::value;
    constexpr int NR = RowCount<N, W, J, REL>::value;
    T rs[VX][NR];
#pragma unroll
    for (int m = 0; m < MID; ++m)
#pragma unroll
        for (int row = 0; row < NR; ++row)
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                T val = T(0);
#pragma unroll
                for (int a = 0; a < W; ++a) val = fma(C.w1[v][a], raw[m][row][v][a], val);
                rs[v][row] = (m == 0) ? (N <= 3 ? val : C.wm[0] * val) : fma(C.wm[m], val, rs[v][row]);
            }
    if (N < 3) {
#pragma unroll
        for (int v = 0; v < VX; ++v) res[0][v] = rs[v][0];
        return;
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int shift = (j == 0) ? 0 : (REL == 2 ? W : REL);
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            T acc = T(0);
#pragma unroll
            for (int b = 0; b < W; ++b) acc = fma(C.w2[j][b], rs[v][b + shift], acc);
            res[j][v] = acc;
        }
    }
}

template <typename T, int N, int W, int VX, int J, int REL>
__device__ __forceinline__ void plane_values(const PencilCtx<T, N, W, VX, J>& C, const T* const (&pv)[VX], int node,
                                             T (&res)[J][VX]) {
    T raw[MidCount<W, N>::value][RowCount<N, W, J, REL>::value][VX][W];
    plane_load<T, N, W, VX, J, REL>(C, pv, node, raw);
    plane_reduce<T, N, W, VX, J, REL>(C, raw, res);
}

// FULL: every output of the register tile exists (no bounds guards in the inner loop).
// PF:   software-pipelined prefetch of the next plane (costs VX*rows*W registers).
//
// The chunk of the last axis is walked run by run: a run is a maximal stretch of consecutive
// grid coordinates that share the same first node i0 (s_run[r] = first coordinate of run r,
// built once per CTA).  Per run: move the window (normally a slide by one node whose loads were
// issued one run earlier), issue the loads of the next node, then a branch-free inner loop over
// the run's outputs.  Keeping the loads in flight across the inner loop only works when that
// loop has a single path — with the window logic inside it the compiler waits for them at the
// loop head (ncu: 66 % of stall samples were long-scoreboard waits on that branch).
template <typename T, int N, int W, int VX, int J, bool NURBS, int REL, bool FULL, bool PFREQ>
__device__ __forceinline__ void walk_pencil(const EvalParams<T>& P, const PencilCtx<T, N, W, VX, J>& C, const int* s_i0,
                                            const T* s_w, const int* s_run, int n_runs, int64_t R, int64_t out_off,
                                            int64_t jstride, const bool (&ok)[J][VX]) {
    constexpr int WB = NURBS ? W : 1;
    constexpr int MID = MidCount<W, N>::value;
    constexpr int NR = RowCount<N, W, J, REL>::value;
    constexpr int RAWREGS = MID * NR * VX * W * (int)(sizeof(T) / 4);
    constexpr bool PF = PFREQ && !NURBS && RAWREGS <= 32;
    const int nN = P.dim[N - 1].n_u;
    for (int c = 0; c < P.n_comp; ++c) {
        const T* __restrict__ uc = P.u + (int64_t)c * P.u_comp_stride;
        // output row pointers (one per j), advanced by R per step; made opaque like the u pointers
        T* op[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            op[j] = P.out + (int64_t)c * P.out_comp_stride + out_off + j * jstride;
            asm volatile("" : "+l"(op[j]));
        }
        // per-output base pointers, made opaque so that the 64-bit component/base offsets are not
        // re-derived for every load (ncu: address arithmetic was > 50 % of the inner loop)
        const T* pu[VX];
        const T* pw[VX];
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            pu[v] = uc + C.off1[v];
            pw[v] = NURBS ? P.wts + C.off1[v] : nullptr;
            asm volatile("" : "+l"(pu[v]));
            if (NURBS) asm volatile("" : "+l"(pw[v]));
        }
        T A[W][J][VX], B[WB][J][VX];
        T pending[PF ? MID : 1][PF ? NR : 1][VX][W];
        auto& pend_raw = reinterpret_cast<T(&)[MID][NR][VX][W]>(pending);
        int cur = -(1 << 30);
        int g = s_run[0];
        for (int r = 0; r < n_runs; ++r) {
            const int g_end = s_run[r + 1];
            const int i0 = s_i0[g];
            const int shift = i0 - cur;
            if (PF && shift == 1) {
                // pending holds node cur + W (issued when the window moved to cur)
#pragma unroll
                for (int a = 0; a + 1 < W; ++a)
#pragma unroll
                    for (int j = 0; j < J; ++j)
#pragma unroll
                        for (int v = 0; v < VX; ++v) A[a][j][v] = A[a + 1][j][v];
                plane_reduce<T, N, W, VX, J, REL>(C, pend_raw, A[W - 1]);
            } else if (shift > 0 && shift < W) {
                for (int s = 0; s < shift; ++s) {   // slide the window one node at a time
#pragma unroll
                    for (int a = 0; a + 1 < W; ++a)
#pragma unroll
                        for (int j = 0; j < J; ++j)
#pragma unroll
                            for (int v = 0; v < VX; ++v) {
                                A[a][j][v] = A[a + 1][j][v];
                                if (NURBS) B[a][j][v] = B[a + 1][j][v];
                            }
                    plane_values<T, N, W, VX, J, REL>(C, pu, cur + s + W, A[W - 1]);
                    if (NURBS) plane_values<T, N, W, VX, J, REL>(C, pw, cur + s + W, B[WB - 1]);
                }
            } else {   // first run of the chunk, or t_eval not sorted / sparser than the nodes
#pragma unroll
                for (int a = 0; a < W; ++a) {
                    plane_values<T, N, W, VX, J, REL>(C, pu, i0 + a, A[a]);
                    if (NURBS) plane_values<T, N, W, VX, J, REL>(C, pw, i0 + a, B[NURBS ? a : 0]);
                }
            }
            cur = i0;
            if constexpr (PF) plane_load<T, N, W, VX, J, REL>(C, pu, min(cur + W, nN - 1), pend_raw);
            // one instruction per warp and run: every lane asks L1 for one line of the plane pf1 nodes ahead
            if (C.pf1 > 0 && C.pf_off >= 0) {
                const int64_t o = C.pf_off + (int64_t)min(cur + W + C.pf1 - (PF ? 0 : 1), nN - 1) * C.sN;
                prefetch_l1(uc + o);
                if (NURBS) prefetch_l1(P.wts + o);
            }
            // branch-free inner loop over the outputs of this run
            for (; g < g_end; ++g) {
                T wl[W];
#pragma unroll
                for (int a = 0; a < W; ++a) wl[a] = s_w[g * W + a];
#pragma unroll
                for (int j = 0; j < J; ++j) {
#pragma unroll
                    for (int v = 0; v < VX; ++v) {
                        T acc = T(0), den = T(0);
#pragma unroll
                        for (int a = 0; a < W; ++a) {
                            acc = fma(wl[a], A[a][j][v], acc);
                            if (NURBS) den = fma(wl[a], B[NURBS ? a : 0][j][v], den);
                        }
                        if (FULL || ok[j][v]) __stcs(op[j] + 32 * v, NURBS ? nurbs_div(acc, den) : acc);
                    }
                    op[j] += R;
                }
            }
        }
    }
}

// One warp = one work item: 32*VX consecutive g_1 x J consecutive g_2 at fixed middle axes,
// over `chunk` coordinates of the last axis.
template <typename T, int N, int W, int VX, int J, bool NURBS, bool FULL, bool PF>
__global__ void __launch_bounds__(PBLOCK, PF ? 4 : ((sizeof(T) == 8 || W >= 3 || N >= 4) ? 4 : 8)) grid_pencil_kernel(const __grid_constant__ EvalParams<T> P, int n_items, int nseg,
                                                             int n2g, int chunk, int l2_warm, int pf1) {
    using Ctx = PencilCtx<T, N, W, VX, J>;
    __shared__ int s_i0[MAXCHUNK];
    __shared__ T s_w[MAXCHUNK * W];
    __shared__ int s_run[MAXCHUNK + 1];
    __shared__ int s_nruns;
    const DimParams<T>& L = P.dim[N - 1];
    // L2 warm-up: the first CTAs to run ask the TMA unit to pull all of u (and the NURBS weights)
    // into L2.  Without it the walk's loads of not-yet-touched u go to HBM behind the output write
    // stream, and their latency (not bandwidth) bounds the kernel (ncu: long scoreboard at the
    // window update even with one-run-ahead software prefetch).
    if (l2_warm && threadIdx.x == 0) {
        constexpr unsigned NPF = 148 * 4;
        constexpr size_t SLICE = 32768;
        const unsigned lin = blockIdx.y * gridDim.x + blockIdx.x;
        if (lin < NPF) {
            const size_t total = (size_t)P.u_comp_stride * P.n_comp * sizeof(T);
            for (size_t off = (size_t)lin * SLICE; off < total; off += (size_t)NPF * SLICE) {
                const uint32_t n = (uint32_t)min(SLICE, total - off) & ~15u;
                if (n) prefetch_l2_bulk(reinterpret_cast<const char*>(P.u) + off, n);
            }
            if (NURBS) {
                const size_t wtotal = (size_t)P.u_comp_stride * sizeof(T);
                for (size_t off = (size_t)lin * SLICE; off < wtotal; off += (size_t)NPF * SLICE) {
                    const uint32_t n = (uint32_t)min(SLICE, wtotal - off) & ~15u;
                    if (n) prefetch_l2_bulk(reinterpret_cast<const char*>(P.wts) + off, n);
                }
            }
        }
    }
    const int g_lo = blockIdx.y * chunk;
    const int g_n = min((int)L.n_eval - g_lo, chunk);
    for (int i = threadIdx.x; i < g_n; i += PBLOCK) s_i0[i] = __ldg(L.tab_i0 + g_lo + i);
    for (int i = threadIdx.x; i < g_n * W; i += PBLOCK) s_w[i] = __ldg(L.tab_w + (int64_t)g_lo * W + i);
    __syncthreads();
    // runs of equal first node: s_run[r] = first coordinate (relative to g_lo) of run r, s_run[n_runs] = g_n
    if (threadIdx.x < 32) {
        int n = 0;
        for (int b0 = 0; b0 < g_n; b0 += 32) {
            const int t = b0 + threadIdx.x;
            const bool start = t < g_n && (t == 0 || s_i0[t] != s_i0[t - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, start);
            if (start) s_run[n + __popc(m & ((1u << threadIdx.x) - 1))] = t;
            n += __popc(m);
        }
        if (threadIdx.x == 0) { s_run[n] = g_n; s_nruns = n; }
    }
    __syncthreads();
    const int n_runs = s_nruns;
    const int item = blockIdx.x * PWARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int lane = threadIdx.x & 31;

    Ctx C;
    const int n1 = (int)P.dim[0].n_eval;
    int q = item / nseg;
    const int g1base = (item - q * nseg) * (32 * VX) + lane;
    bool ok[J][VX];
    bool ok1[VX];
#pragma unroll
    for (int v = 0; v < VX; ++v) {
        const int g1 = g1base + 32 * v;
        ok1[v] = g1 < n1;
        const int g1c = min(g1, n1 - 1);
        C.off1[v] = __ldg(P.dim[0].tab_i0 + g1c);
#pragma unroll
        for (int a = 0; a < W; ++a) C.w1[v][a] = __ldg(P.dim[0].tab_w + g1c * W + a);
    }
    int64_t out_off = g1base;
    int base = 0;
    int rel = 0;
    C.row0 = 0;
    C.rel_rows = 0;
    C.s1 = (N >= 3) ? (int)P.dim[1].u_stride : 0;
    if (N >= 3) {
        const int n2 = (int)P.dim[1].n_eval;
        const int rest = q / n2g;
        const int g2 = (q - rest * n2g) * J;
        q = rest;
        out_off += P.dim[1].out_stride * g2;
        int i02[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int g2c = min(g2 + j, n2 - 1);
            i02[j] = __ldg(P.dim[1].tab_i0 + g2c);
#pragma unroll
            for (int b = 0; b < W; ++b) C.w2[j][b] = __ldg(P.dim[1].tab_w + g2c * W + b);
#pragma unroll
            for (int v = 0; v < VX; ++v) ok[j][v] = ok1[v] && (g2 + j < n2);
        }
        C.row0 = i02[0] * C.s1;
        if (J == 2) {
            rel = i02[J - 1] - i02[0];
            C.rel_rows = rel;
        }
    } else {
#pragma unroll
        for (int v = 0; v < VX; ++v) ok[0][v] = ok1[v];
    }
    // middle axes 3..N-1
    {
        int gm[2] = {0, 0};
#pragma unroll
        for (int d = 2; d < N - 1; ++d) {
            const int ne = (int)P.dim[d].n_eval;
            const int qq = q / ne;
            gm[d - 2] = q - qq * ne;
            q = qq;
            out_off += P.dim[d].out_stride * gm[d - 2];
            base += __ldg(P.dim[d].tab_i0 + gm[d - 2]) * (int)P.dim[d].u_stride;
        }
        // "point" axes (stencil width 1, e.g. Constant dimensions between BSpline ones): they sit after
        // the N stencil axes in P.dim (the host permutes them there) and only select a slice of u
        for (int e = N; e < P.n_in; ++e) {
            const int ne = (int)P.dim[e].n_eval;
            const int qq = q / ne;
            const int g = q - qq * ne;
            q = qq;
            out_off += P.dim[e].out_stride * g;
            base += __ldg(P.dim[e].tab_i0 + g) * (int)P.dim[e].u_stride;
        }
#pragma unroll
        for (int m = 0; m < Ctx::MID; ++m) {
            if (N <= 3) { C.wm[m] = T(1); C.om[m] = 0; }
            else if (N == 4) {
                C.wm[m] = __ldg(P.dim[2].tab_w + gm[0] * W + m);
                C.om[m] = m * (int)P.dim[2].u_stride;
            } else {
                C.wm[m] = __ldg(P.dim[2].tab_w + gm[0] * W + (m % W)) * __ldg(P.dim[3].tab_w + gm[1] * W + (m / W));
                C.om[m] = (m % W) * (int)P.dim[2].u_stride + (m / W) * (int)P.dim[3].u_stride;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VX; ++v) C.off1[v] += base;
    C.sN = (int)L.u_stride;
    C.pf1 = pf1;
    C.pf_off = -1;
    constexpr int PF_NRMAX = Ctx::MID * ((N < 3) ? 1 : (J == 1 ? W : 2 * W));
    if (pf1 > 0 && PF_NRMAX <= 32) {
        // lane -> (middle stencil point, row, 128-byte segment) of the warp's footprint in one plane of u: rows
        // as plane_load walks them, segments from lane 0's first node of output 0 up to lane 31's last node of
        // output VX-1
        constexpr int SEGS = 32 / PF_NRMAX;
        constexpr int EL = 128 / (int)sizeof(T);
        const int first = __shfl_sync(0xffffffffu, C.off1[0], 0);
        const int last = __shfl_sync(0xffffffffu, C.off1[VX - 1], 31) + W - 1;
        const int mr = lane / SEGS, seg = lane - mr * SEGS;
        const int m = mr % Ctx::MID, row = mr / Ctx::MID;
        const bool unrel = !(rel == 0 || rel == 1);
        const int nrows = (N < 3) ? 1 : (J == 1 ? W : (unrel ? 2 * W : W + rel));
        if (row < nrows && first + seg * EL <= last) {
            const int ro = (row < W || !unrel) ? row : rel + row - W;
            int omm = 0;
#pragma unroll
            for (int q2 = 0; q2 < Ctx::MID; ++q2) if (q2 == m) omm = C.om[q2];
            const int64_t o = (int64_t)first + omm + C.row0 + (int64_t)ro * C.s1 + seg * EL;
            if (o + (int64_t)(P.dim[N - 1].n_u - 1) * C.sN < P.u_comp_stride) C.pf_off = (int)o;
        }
        // the chunk's first window and the planes the walk's own prefetch will not ask for
        if (C.pf_off >= 0) {
            const int nN = P.dim[N - 1].n_u, i0 = s_i0[0];
            for (int a = 0; a < W + pf1; ++a) {
                const int64_t o = C.pf_off + (int64_t)min(i0 + a, nN - 1) * C.sN;
                prefetch_l1(P.u + o);
                if (NURBS) prefetch_l1(P.wts + o);
            }
        }
    }
    const int64_t R = L.out_stride;              // points per last-axis coordinate
    out_off += (int64_t)g_lo * R;
    const int64_t jstride = (N >= 3) ? P.dim[1].out_stride : 0;

    if (J == 1 || N < 3) walk_pencil<T, N, W, VX, J, NURBS, 0, FULL, PF>(P, C, s_i0, s_w, s_run, n_runs, R, out_off, jstride, ok);
    else if (rel == 0) walk_pencil<T, N, W, VX, J, NURBS, 0, FULL, PF>(P, C, s_i0, s_w, s_run, n_runs, R, out_off, jstride, ok);
    else if (rel == 1) walk_pencil<T, N, W, VX, J, NURBS, 1, FULL, PF>(P, C, s_i0, s_w, s_run, n_runs, R, out_off, jstride, ok);
    else walk_pencil<T, N, W, VX, J, NURBS, 2, FULL, PF>(P, C, s_i0, s_w, s_run, n_runs, R, out_off, jstride, ok);
}

template <typename T, int N, int W, int VX, int J>
cudaError_t launch_pencil_cfg(const EvalParams<T>& P, cudaStream_t s) {
    const int n1 = (int)P.dim[0].n_eval;
    const int nseg = (n1 + 32 * VX - 1) / (32 * VX);
    const int n2g = (N >= 3) ? ((int)P.dim[1].n_eval + J - 1) / J : 1;
    int64_t items = (int64_t)nseg * n2g;
    for (int d = 2; d < N - 1; ++d) items *= P.dim[d].n_eval;
    for (int e = N; e < P.n_in; ++e) items *= P.dim[e].n_eval;
    const int n_last = (int)P.dim[N - 1].n_eval;
    const int bx = (int)((items + PWARPS - 1) / PWARPS);
    // >= ~6 waves of CTAs (short tail) while keeping pencils long enough to amortise the window start-up
    // measured on config 2 with the L1 prefetch: 64 -> 0.114 ms, 52 -> 0.116, 32 -> 0.118, 16 -> 0.131 (window
    // start-up), 86 -> 0.125, 128 -> 0.134 (too few CTAs: 1.4 waves)
    int chunk = 64;
    {
        // ... unless 64 leaves a visibly emptier last wave than a shorter chunk (CTAs per SM as in the launch bounds):
        // the longest of 64, 56, 48, 40, 32 that fills its last wave to >= 94 %, else the best-filled one.  Config 2
        // (592 CTA slots): 64 -> 3.46 waves 0.114 ms, 48 -> 4.76 waves 0.108 ms, 32 -> 6.9 waves 0.111 ms.
        // No window to start up (W = 1): 32.
        const bool pfk = !P.wts && sizeof(T) == 4 && W == 2;
        const double slots = 148.0 * ((pfk || sizeof(T) == 8 || W >= 3 || N >= 4) ? 4 : 8);
        auto fill = [&](int ch) {
            const double waves = (double)bx * ((n_last + ch - 1) / ch) / slots;
            return waves / ceil(waves);
        };
        if (W == 1) chunk = 32;
        else {
            double best = -1.0;
            for (int ch : {64, 56, 48, 40, 32}) {
                const double f = fill(ch);
                if (f >= 0.94) { chunk = ch; break; }
                if (f > best + 0.02) { best = f; chunk = ch; }
            }
        }
    }
    while (chunk > 8 && (int64_t)bx * ((n_last + chunk - 1) / chunk) < 148 * 5 * 2) chunk >>= 1;
    const Knobs& K = knobs_of(P.knobs);
    if (K.chunk > 0) chunk = K.chunk;
    chunk = min(MAXCHUNK, max(1, chunk));
    dim3 grid((unsigned)bx, (unsigned)((n_last + chunk - 1) / chunk));
    // L1 prefetch distance, measured: config 2 0 -> 0.133 ms, 1..3 -> 0.122 ms, 6 -> 0.128 ms; every W = 2 / 3 case
    // of scripts/sanity_perf.py gains 2-16 %, the W = 4 cases (long runs, 3-4x upsampling) lose 3-6 %
    // W = 4: the L1 prefetch pays when the walk changes node often (<= ~2.5 outputs per node of the last axis:
    // tricubic F32 256^3 -> 512^3 520 -> 588 Gpoints/s) and costs 6-8 % on long runs (128^3 -> 384^3: 810 -> 763)
    const bool dense_walk = (double)n_last <= 2.5 * (double)P.dim[N - 1].n_u;
    const int pf1_default = W <= 3 ? 4 : (dense_walk ? 2 : 0);   // 4 CTAs/SM: distance 4..8 beats 2 by 2 % on config 2 (L1 has the room)
    // warm L2 with u when it (plus the NURBS weights) fits comfortably and the pointers allow bulk prefetch; with the
    // L1 prefetch running ahead of the walk the warm-up only competes with it (config 2: 0.116 with, 0.114 ms without)
    const size_t u_bytes = (size_t)P.u_comp_stride * (P.n_comp + (P.wts ? 1 : 0)) * sizeof(T);
    int l2_warm = u_bytes >= (1u << 20) && u_bytes <= (96u << 20) && reinterpret_cast<uintptr_t>(P.u) % 16 == 0 &&
                  (!P.wts || reinterpret_cast<uintptr_t>(P.wts) % 16 == 0);
    l2_warm = K.l2_warm >= 0 ? K.l2_warm : (l2_warm && pf1_default == 0);
    const bool full = n1 % (32 * VX) == 0 && (N < 3 || P.dim[1].n_eval % J == 0);
    const bool pf = K.pf != 0;
    const int pf1 = K.pf1 >= 0 ? K.pf1 : pf1_default;
    if (P.wts) {
        if (full) grid_pencil_kernel<T, N, W, VX, J, true, true, false><<<grid, PBLOCK, 0, s>>>(P, (int)items, nseg, n2g, chunk, l2_warm, pf1);
        else grid_pencil_kernel<T, N, W, VX, J, true, false, false><<<grid, PBLOCK, 0, s>>>(P, (int)items, nseg, n2g, chunk, l2_warm, pf1);
    } else if (sizeof(T) == 4 && W == 2 && pf) {
        if (full) grid_pencil_kernel<T, N, W, VX, J, false, true, true><<<grid, PBLOCK, 0, s>>>(P, (int)items, nseg, n2g, chunk, l2_warm, pf1);
        else grid_pencil_kernel<T, N, W, VX, J, false, false, true><<<grid, PBLOCK, 0, s>>>(P, (int)items, nseg, n2g, chunk, l2_warm, pf1);
    } else {
        if (full) grid_pencil_kernel<T, N, W, VX, J, false, true, false><<<grid, PBLOCK, 0, s>>>(P, (int)items, nseg, n2g, chunk, l2_warm, pf1);
        else grid_pencil_kernel<T, N, W, VX, J, false, false, false><<<grid, PBLOCK, 0, s>>>(P, (int)items, nseg, n2g, chunk, l2_warm, pf1);
    }
    return cudaGetLastError();
}

// Register budget: the compiler keeps all VX * rows * W loads of one plane in flight.
template <typename T, int W>
struct PickVX {
    static constexpr int per = 2 * W * W * (int)(sizeof(T) / 4);
    static constexpr int value = (4 * per <= 64) ? 4 : (2 * per <= 64) ? 2 : 1;
};

template <typename T, int N, int W>
cudaError_t launch_pencil(const EvalParams<T>& P, cudaStream_t s) {
    constexpr int VXF = PickVX<T, W>::value;
    constexpr int J = (N >= 3) ? 2 : 1;
    const int n1 = (int)P.dim[0].n_eval;
    if (VXF > 1 && n1 > 32 * (VXF / 2)) return launch_pencil_cfg<T, N, W, VXF, J>(P, s);
    return launch_pencil_cfg<T, N, W, 1, J>(P, s);
}

// Dispatch for NLO <= N_in <= NHI (the instantiations are spread over several translation
// units to keep the build parallel).
template <typename T, int NLO, int NHI>
bool try_grid_range(const EvalParams<T>& P0, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    const int NT = P0.n_in;
    if (NT < 2) return false;
    // Stencil axes (width W) keep their order in front; axes of width 1 (Constant) that lie strictly
    // between axis 2 and the last axis become "point" axes behind them.
    int W = 1;
    for (int d = 0; d < NT; ++d) W = max(W, P0.dim[d].width);
    if (W < 1 || W > 4) return false;   // W = 1: all-Constant dimensions (nearest-node lookup)
    EvalParams<T> P = P0;
    int ns = 0, np = 0;
    if (P0.n_stencil > 0) {   // axes already classified and ordered by the caller
        ns = P0.n_stencil;
        np = NT - ns;
    } else {
        int point[MAXN];
        for (int d = 0; d < NT; ++d) {
            if (P0.dim[d].width == W) P.dim[ns++] = P0.dim[d];
            else if (P0.dim[d].width == 1 && d >= 2 && d < NT - 1) point[np++] = d;
            else return false;
        }
        if (P0.dim[0].width != W || P0.dim[NT - 1].width != W || (ns >= 3 && P0.dim[1].width != W)) return false;
        for (int e = 0; e < np; ++e) P.dim[ns + e] = P0.dim[point[e]];
    }
    const int N = ns;
    if (N < NLO || N > NHI || N < 2 || N > 5) return false;
    int64_t total = 1;
    for (int d = 0; d < NT; ++d) {
        if (P.dim[d].n_eval >= (1 << 30)) return false;
        total *= P.dim[d].n_eval;
    }
    if (P.u_comp_stride >= (1LL << 31)) return false;
    if (total / P.dim[N - 1].n_eval >= (1LL << 31) || P.dim[N - 1].n_eval >= 65535) return false;
    *launches = 1;
#define DIND_PENCIL(NN, WW)                                           \
    if (NN >= NLO && NN <= NHI && N == NN && W == WW) {               \
        *err = launch_pencil<T, (NN >= NLO && NN <= NHI) ? NN : NLO, WW>(P, s); \
        *name = np ? "grid_pencil<N=" #NN "+point axes,W=" #WW ">" : "grid_pencil<N=" #NN ",W=" #WW ">"; \
        return true;                                                  \
    }
    DIND_PENCIL(2, 1) DIND_PENCIL(3, 1) DIND_PENCIL(4, 1) DIND_PENCIL(5, 1)
    DIND_PENCIL(2, 2) DIND_PENCIL(2, 3) DIND_PENCIL(2, 4)
    DIND_PENCIL(3, 2) DIND_PENCIL(3, 3) DIND_PENCIL(3, 4)
    DIND_PENCIL(4, 2) DIND_PENCIL(4, 3) DIND_PENCIL(4, 4)
    DIND_PENCIL(5, 2)
#undef DIND_PENCIL
    return false;
}

}  // namespace

}  // namespace dind
