// dind_fast_grid_split.cu — eval_grid! for 4-D B-spline / NURBS grids as TWO separable stages.
//
// The pencil kernel (dind_fast_grid.cuh) contracts axes 1..N-1 inside every thread for each new
// plane of the last axis: for BASELINE config 4 (4-D, degree 2, NURBS) that is 27 nodes x 2 arrays
// per thread and plane, re-gathered through L1 by every (g_2, g_3) item — ncu: L1 load wavefronts
// 86 %, 0.19 of the HBM roofline.  A tensor-product stencil factorises, so the evaluation is split:
//
//   stage 1  T(g_1, g_2, i_3, i_4) = sum_{a1,a2} w_1 w_2 u(i_1+a1, i_2+a2, i_3, i_4)
//            = the 2-D pencil kernel over axes (1, 2) with axes 3 and 4 as identity "point" axes
//            (NURBS: once for u*w, once for w);
//   stage 2  out(b, g_3, g_4) = sum_{a3,a4} w_3 w_4 T(b, i_3+a3, i_4+a4),  b = (g_1, g_2) contiguous
//            = grid_batch_kernel below: lanes along the batch index (every load and store fully
//            coalesced, all stencil indices and weights warp-uniform), two rows of g_3 per thread,
//            a sliding window of W planes along axis 4; NURBS divides at the end.
//
// FMAs per output drop from ~30 to ~6 (x2 for NURBS) and nothing is gathered; the price is one
// round trip of T through HBM/L2 (config 4: 2 x 537 MB written and read once).
#include <algorithm>
#include <cstring>

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int SB_BLOCK = 128;
constexpr int SB_WARPS = SB_BLOCK / 32;
constexpr int SB_MAXCHUNK = 64;
constexpr int SB_VB = 2;
#ifndef SB_PF
#define SB_PF 1
#endif

template <typename T>
struct BatchParams {
    const T* tn;          // numerator (or the only) stage-1 result, (B, n3, n4)
    const T* td;          // denominator (NURBS) or null
    T* out;               // (B, g3, g4)
    int64_t B;
    int n3, n4, g3, g4;
    const int32_t* i0_3;
    const T* w_3;
    const int32_t* i0_4;
    const T* w_4;
    int chunk;
    int n_bseg, n_items;
    int pf;               // planes of axis 4 the L2 prefetch runs ahead of the walk
};

__global__ void iota_kernel(int32_t* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// One plane of the window: rows of T for the thread's 2 x VB outputs, contracted along axis 3 into window
// slot PH.  The slot is a template parameter so that the window never moves between registers: the walk
// switches (warp-uniformly) on the slot and rotates the three or four axis-4 WEIGHTS instead.
// CO: T is being produced by other CTAs of the SAME launch (fused kernel below): read it with ordinary loads, not
// through the non-coherent path.
template <typename T, bool CO>
__device__ __forceinline__ T ld_t(const T* p) {
    if constexpr (CO) return *p;
    else return __ldg(p);
}

// two adjacent batch columns as one 8- / 16-byte load or store
template <typename T> struct Pair2;
template <> struct Pair2<float> { using type = float2; };
template <> struct Pair2<double> { using type = double2; };

// VEC: the thread's two batch columns are ADJACENT (b, b + 1; B even): one vector load per row of T instead of two
// scalar ones — half the loads, prefetches and address arithmetic of the walk (ncu: 22 of the 58 instructions per output
// were integer address arithmetic, the kernel is bound by instruction issue).
template <typename T, int W, int W4, bool NURBS, int REL, int PH, bool CO = false, bool VEC = false>
__device__ __forceinline__ void batch_plane(const T* const (&bn)[SB_VB], int64_t pl, int64_t dd, int64_t stepB, int64_t rel_b,
                                            int64_t pf_off, bool pf_ok, const T (&w3)[2][W], T (&wn)[W4][2][SB_VB],
                                            T (&wd)[NURBS ? W4 : 1][2][SB_VB]) {
    constexpr int VB = SB_VB;
    constexpr int ROWS = REL == 2 ? 2 * W : W + REL;
    T rn[VB][ROWS], rd[VB][NURBS ? ROWS : 1];
    if constexpr (VEC) {
        static_assert(SB_VB == 2 && !CO, "pair loads");
        using P2 = typename Pair2<T>::type;
        const T* q = bn[0] + pl;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const P2 a = __ldg(reinterpret_cast<const P2*>(q));
            rn[0][r] = a.x;
            rn[1][r] = a.y;
            if (NURBS) {
                const P2 d = __ldg(reinterpret_cast<const P2*>(q + dd));
                rd[0][r] = d.x;
                rd[1][r] = d.y;
            }
            if (pf_ok) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q + pf_off));
                if (NURBS) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + dd + pf_off));
            }
            q += (REL == 2 && r == W - 1) ? rel_b : stepB;
        }
    } else {
#pragma unroll
        for (int v = 0; v < VB; ++v) {
            const T* q = bn[v] + pl;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                rn[v][r] = ld_t<T, CO>(q);
                if (NURBS) rd[v][r] = ld_t<T, CO>(q + dd);
                // the walk is sequential along axis 4: ask for the same rows of the next plane now
                if (pf_ok) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + pf_off));
                    if (NURBS) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + dd + pf_off));
                }
                q += (REL == 2 && r == W - 1) ? rel_b : stepB;
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VB; ++v)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r0 = j == 0 ? 0 : (REL == 2 ? W : REL);
            T an = T(0), ad = T(0);
#pragma unroll
            for (int a = 0; a < W; ++a) {
                an = fma(w3[j][a], rn[v][r0 + a], an);
                if (NURBS) ad = fma(w3[j][a], rd[v][r0 + a], ad);
            }
            wn[PH][j][v] = an;
            if (NURBS) wd[PH][j][v] = ad;
        }
}

// REL = 0 / 1: the two g_3 rows of the thread start at the same node / one node apart and share
// their W + REL rows of T; REL = 2: unrelated rows (down-sampling, unsorted t_eval), loaded separately.
template <typename T, int W, int W4, bool NURBS, int REL, bool CO = false, bool VEC = false>
__device__ __forceinline__ void batch_walk(const BatchParams<T>& Q, const int* s_i0, const T* s_w, int g_lo, int g_n,
                                           const int64_t (&boff)[SB_VB], const bool (&okb)[SB_VB], int g3a, bool ok1,
                                           int i3a, int i3b, const T (&w3)[2][W]) {
    constexpr int VB = SB_VB;
    T wn[W4][2][VB];
    T wd[NURBS ? W4 : 1][2][VB];
#pragma unroll
    for (int a = 0; a < W4; ++a)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int v = 0; v < VB; ++v) {
                wn[a][j][v] = T(0);
                if (NURBS) wd[a][j][v] = T(0);
            }
    const int64_t Bn3 = Q.B * Q.n3;
    const T* bn[VB];
#pragma unroll
    for (int v = 0; v < VB; ++v) bn[v] = Q.tn + boff[v] + Q.B * (int64_t)i3a;
    const int64_t dd = NURBS ? Q.td - Q.tn : 0;                       // same offsets in both stage-1 arrays
    const int64_t rel_b = Q.B * (int64_t)(i3b - i3a - (W - 1));       // REL == 2: jump from row i3a + W - 1 to row i3b
    const int64_t pf_off = Q.pf * Bn3;   // L2 prefetch distance in planes
    int cur = -(1 << 30);
    int phase = 0;            // window slot that receives the next plane; plane cur + a lives in slot (phase + a) % W
    int rot[W4];               // rot[s] = (s - phase) mod W: the axis-4 weight that belongs to slot s
#pragma unroll
    for (int sl = 0; sl < W4; ++sl) rot[sl] = sl;
    const int64_t o_plane = Q.B * Q.g3;
    T* op[2][VB];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int v = 0; v < VB; ++v) op[j][v] = Q.out + Q.B * ((int64_t)(g3a + j) + (int64_t)Q.g3 * g_lo) + boff[v];
    for (int g = 0; g < g_n; ++g) {
        const int i4 = s_i0[g];
        if (i4 != cur) {   // warp-uniform
            const int shift = (i4 > cur && i4 - cur < W4) ? i4 - cur : W4;
            for (int s = 0; s < shift; ++s) {
                const int pidx = i4 + (W4 - shift) + s;
                const bool pf_ok = pidx + Q.pf < Q.n4;
                const int64_t pl = Bn3 * (int64_t)pidx;
                if (W4 == 1 || phase == 0) batch_plane<T, W, W4, NURBS, REL, 0, CO, VEC>(bn, pl, dd, Q.B, rel_b, pf_off, pf_ok, w3, wn, wd);
                if constexpr (W4 > 1) {
                    if (phase == 1) batch_plane<T, W, W4, NURBS, REL, 1, CO, VEC>(bn, pl, dd, Q.B, rel_b, pf_off, pf_ok, w3, wn, wd);
                }
                if constexpr (W4 > 2) {
                    if (phase == 2) batch_plane<T, W, W4, NURBS, REL, 2, CO, VEC>(bn, pl, dd, Q.B, rel_b, pf_off, pf_ok, w3, wn, wd);
                }
                if constexpr (W4 > 3) {
                    if (phase == 3) batch_plane<T, W, W4, NURBS, REL, 3, CO, VEC>(bn, pl, dd, Q.B, rel_b, pf_off, pf_ok, w3, wn, wd);
                }
                phase = phase + 1 == W4 ? 0 : phase + 1;
            }
            cur = i4;
#pragma unroll
            for (int sl = 0; sl < W4; ++sl) rot[sl] = sl - phase + (sl < phase ? W4 : 0);
        }
        T w4[W4];
#pragma unroll
        for (int sl = 0; sl < W4; ++sl) w4[sl] = s_w[g * W4 + rot[sl]];
        T qn[2][VB], qd[2][VB];
        bool normal = true;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int v = 0; v < VB; ++v) {
                T an = T(0), ad = T(1);
                if (NURBS) ad = T(0);
#pragma unroll
                for (int sl = 0; sl < W4; ++sl) {
                    an = fma(w4[sl], wn[sl][j][v], an);
                    if (NURBS) ad = fma(w4[sl], wd[sl][j][v], ad);
                }
                qn[j][v] = an;
                qd[j][v] = ad;
                if (NURBS) normal = normal && nurbs_den_is_normal(ad);
            }
        // one (practically never taken) branch per step instead of one per quotient
        if (NURBS && !normal) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int v = 0; v < VB; ++v) qn[j][v] = qn[j][v] / qd[j][v];
        } else if (NURBS) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int v = 0; v < VB; ++v) qn[j][v] = nurbs_div_normal(qn[j][v], qd[j][v]);
        }
        if constexpr (VEC) {
            using P2 = typename Pair2<T>::type;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                P2 pr;
                pr.x = qn[j][0];
                pr.y = qn[j][1];
                if (okb[0] && (j == 0 || ok1)) __stcs(reinterpret_cast<P2*>(op[j][0]), pr);
                op[j][0] += o_plane;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int v = 0; v < VB; ++v) {
                    if (okb[v] && (j == 0 || ok1)) __stcs(op[j][v], qn[j][v]);
                    op[j][v] += o_plane;
                }
        }
    }
}

// One CTA's share of stage 2: block (bx, by) of NT threads; bseg0 = first batch segment of the B-tile this block
// works in (0 for the stand-alone kernel).  s_i0 / s_w: shared-memory staging of the axis-4 table slice.
// NT threads cooperate: a CTA (tid = threadIdx.x) or, in the fused kernel, ONE warp (NT = 32, tid = lane)
template <int NT>
__device__ __forceinline__ void group_sync() {
    if constexpr (NT == 32) __syncwarp();
    else __syncthreads();
}

template <typename T, int W, int W4, bool NURBS, int NT, bool CO, bool VEC = false>
__device__ __forceinline__ void batch_body(const BatchParams<T>& Q, int bx, int by, int bseg0, int* s_i0, T* s_w, int tid) {
    const int g_lo = by * Q.chunk;
    const int g_n = min(Q.g4 - g_lo, Q.chunk);
    for (int i = tid; i < g_n; i += NT) s_i0[i] = __ldg(Q.i0_4 + g_lo + i);
    for (int i = tid; i < g_n * W4; i += NT) s_w[i] = __ldg(Q.w_4 + (int64_t)g_lo * W4 + i);
    group_sync<NT>();
    const int item = bx * (NT / 32) + (tid >> 5);
    if (item >= Q.n_items) return;
    const int lane = tid & 31;
    const int g3p = item / Q.n_bseg;
    const int bseg = bseg0 + item - g3p * Q.n_bseg;
    int64_t boff[SB_VB];
    bool okb[SB_VB];
#pragma unroll
    for (int v = 0; v < SB_VB; ++v) {
        // VEC: columns 2 * lane, 2 * lane + 1 of the segment (B is even: a pair is in or out as a whole)
        const int64_t b = (int64_t)bseg * (32 * SB_VB) + (VEC ? SB_VB * lane + v : lane + 32 * v);
        okb[v] = b < Q.B;
        boff[v] = okb[v] ? b : Q.B - (VEC ? SB_VB - v : 1);
    }
    const int g3a = g3p * 2;
    const bool ok1 = g3a + 1 < Q.g3;
    const int g3b = ok1 ? g3a + 1 : g3a;
    const int i3a = __ldg(Q.i0_3 + g3a), i3b = __ldg(Q.i0_3 + g3b);
    T w3[2][W];
#pragma unroll
    for (int a = 0; a < W; ++a) {
        w3[0][a] = __ldg(Q.w_3 + g3a * W + a);
        w3[1][a] = __ldg(Q.w_3 + g3b * W + a);
    }
    const int rel = i3b - i3a;
    if (rel == 0) batch_walk<T, W, W4, NURBS, 0, CO, VEC>(Q, s_i0, s_w, g_lo, g_n, boff, okb, g3a, ok1, i3a, i3b, w3);
    else if (rel == 1) batch_walk<T, W, W4, NURBS, 1, CO, VEC>(Q, s_i0, s_w, g_lo, g_n, boff, okb, g3a, ok1, i3a, i3b, w3);
    else batch_walk<T, W, W4, NURBS, 2, CO, VEC>(Q, s_i0, s_w, g_lo, g_n, boff, okb, g3a, ok1, i3a, i3b, w3);
}

template <typename T, int W, int W4, bool NURBS, bool VEC>
__global__ void __launch_bounds__(SB_BLOCK, 4) grid_batch_kernel(const __grid_constant__ BatchParams<T> Q) {
    __shared__ int s_i0[SB_MAXCHUNK];
    __shared__ T s_w[SB_MAXCHUNK * W4];
    batch_body<T, W, W4, NURBS, SB_BLOCK, false, VEC>(Q, blockIdx.x, blockIdx.y, 0, s_i0, s_w, threadIdx.x);
}

// ---- stage 1 with the (n_1, n_2) slice of u in shared memory ----------------------------------------
// One CTA per (i_3, i_4): the slice — contiguous in u — is copied to shared memory once (for NURBS the
// slices of u*w and of w side by side), then every thread owns one g_1 and a chunk of g_2 and walks it
// with a sliding window of W plane values (3 conflict-free LDS per new node of axis 2).  Reads of u and
// writes of T are both fully coalesced; nothing is gathered from global memory.
constexpr int S1_BLOCK = 256;
constexpr int S1_CHUNK = 32;

template <typename T>
struct SliceParams {
    const T* a0;          // u (or u*w), slices of n1*n2 elements, slice s at a0 + s * slice_stride
    const T* a1;          // second array (NURBS weights) or null
    const T* premul;      // NURBS weights to multiply a0 by while it is staged (a0 is the raw u), or null
    T* t0;                // (g1, g2, slices)
    T* t1;
    int n1, n2, g1, g2;
    int64_t slice_stride;
    const int32_t* i0_1;
    const T* w_1;
    const int32_t* i0_2;
    const T* w_2;
};

// One slice (i_3, i_4) of stage 1: outputs g_2 in [g2_lo, g2_hi), `sub` consecutive outputs per thread, from the
// rows [r_lo, r_lo + r_n) of the slice staged in shared memory (the whole slice and axis for the stand-alone kernel;
// one thin B-tile and only the rows it touches for the fused kernel).
template <typename T, int W, int NARR, int NT>
__device__ __forceinline__ void slice_body(const SliceParams<T>& Q, int64_t sl, int g2_lo, int g2_hi, int sub, int r_lo, int r_n, T* S,
                                           int tid) {
    const int nn = Q.n1 * r_n;
    {
        const T* src0 = Q.a0 + sl * Q.slice_stride + (int64_t)r_lo * Q.n1;
        const T* srcw = (NARR == 2 ? Q.a1 : Q.premul) + sl * Q.slice_stride + (int64_t)r_lo * Q.n1;
        if (NARR == 2 || Q.premul) {
            // NURBS numerator u*w (src/interpolation_methods.jl:124-131) formed here instead of by a separate pass
            // four elements per thread and step, all eight loads issued before the first shared-memory store
            constexpr int U = 4;
            int i = tid;
            for (; i + (U - 1) * NT < nn; i += U * NT) {
                T uv[U], wv[U];
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    wv[q] = __ldg(srcw + i + q * NT);
                    uv[q] = __ldcs(src0 + i + q * NT);
                }
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    S[i + q * NT] = Q.premul ? uv[q] * wv[q] : uv[q];
                    if (NARR == 2) S[nn + i + q * NT] = wv[q];
                }
            }
            for (; i < nn; i += NT) {
                const T wv = __ldg(srcw + i);
                S[i] = Q.premul ? __ldcs(src0 + i) * wv : __ldcs(src0 + i);
                if (NARR == 2) S[nn + i] = wv;
            }
        } else {
            for (int i = tid; i < nn; i += NT) S[i] = __ldcs(src0 + i);
        }
    }
    group_sync<NT>();
    const int n_sub = (g2_hi - g2_lo + sub - 1) / sub;
    const int n_tasks = Q.g1 * n_sub;
    const int64_t out_slice = (int64_t)Q.g1 * Q.g2;
    for (int task = tid; task < n_tasks; task += NT) {
        const int ch = task / Q.g1;
        const int g1 = task - ch * Q.g1;
        const int i1 = __ldg(Q.i0_1 + g1);
        T w1[W];
#pragma unroll
        for (int a = 0; a < W; ++a) w1[a] = __ldg(Q.w_1 + g1 * W + a);
        T win[NARR][W];
#pragma unroll
        for (int k = 0; k < NARR; ++k)
#pragma unroll
            for (int a = 0; a < W; ++a) win[k][a] = T(0);
        int cur = -(1 << 30);
        const int g_lo = g2_lo + ch * sub, g_hi = min(g2_hi, g_lo + sub);
        T* o0 = Q.t0 + sl * out_slice + g1;
        T* o1 = NARR == 2 ? Q.t1 + sl * out_slice + g1 : nullptr;
        // axis-2 table entry of the NEXT output is requested one step ahead: read where it is used, the
        // L1 latency of i0_2[g] and w_2[g] sat in every step's dependent chain (ncu source view: 36 % of the
        // kernel's stall samples)
        int i2n = __ldg(Q.i0_2 + g_lo);
        T w2n[W];
#pragma unroll
        for (int a = 0; a < W; ++a) w2n[a] = __ldg(Q.w_2 + g_lo * W + a);
        for (int g = g_lo; g < g_hi; ++g) {
            const int i2 = i2n;
            T w2[W];
#pragma unroll
            for (int a = 0; a < W; ++a) w2[a] = w2n[a];
            {
                const int gn = min(g + 1, g_hi - 1);
                i2n = __ldg(Q.i0_2 + gn);
#pragma unroll
                for (int a = 0; a < W; ++a) w2n[a] = __ldg(Q.w_2 + gn * W + a);
            }
            if (i2 != cur) {
                const int shift = (i2 > cur && i2 - cur < W) ? i2 - cur : W;
                for (int s = 0; s < shift; ++s) {
                    const T* row = S + i1 + Q.n1 * (i2 + (W - shift) + s - r_lo);
#pragma unroll
                    for (int k = 0; k < NARR; ++k) {
#pragma unroll
                        for (int a = 0; a + 1 < W; ++a) win[k][a] = win[k][a + 1];
                        T acc = T(0);
#pragma unroll
                        for (int a = 0; a < W; ++a) acc = fma(w1[a], row[k * nn + a], acc);
                        win[k][W - 1] = acc;
                    }
                }
                cur = i2;
            }
#pragma unroll
            for (int k = 0; k < NARR; ++k) {
                T acc = T(0);
#pragma unroll
                for (int a = 0; a < W; ++a) acc = fma(w2[a], win[k][a], acc);
                (k == 0 ? o0 : o1)[(int64_t)g * Q.g1] = acc;
            }
        }
    }
}

template <typename T, int W, int NARR>
__global__ void __launch_bounds__(S1_BLOCK) grid_slice_kernel(const __grid_constant__ SliceParams<T> Q) {
    extern __shared__ __align__(16) unsigned char s1_smem[];
    slice_body<T, W, NARR, S1_BLOCK>(Q, blockIdx.x, 0, Q.g2, S1_CHUNK, 0, Q.n2, reinterpret_cast<T*>(s1_smem), threadIdx.x);
}

// ---- both stages in ONE persistent launch (opt-in: DIND_PERSIST=1) -------------------------------------------
// Run as two kernels, T (config 4: 2 x 537 MB) is written to HBM by stage 1 and read back by stage 2: 2.1 GB of the
// 4.7 GB the evaluation moves, for 2.4 GB algorithmic.  Alternating the two kernels over L2-sized tiles lost to the
// per-launch ramp-up and tail.  Here one grid of resident CTAs drains two ordered task queues,
// warp by warp:
//   stage-2 items (B-tile h, chunk c of the last axis, 64 batch columns x 2 rows of g_3) — the product;
//   stage-1 tasks (B-tile h, slice (i_3, i_4)), in plane order i_4 — the supply.
// A warp takes the next stage-2 item, looks up the planes i_4 its chunk needs and, while one of them is incomplete,
// executes stage-1 tasks itself in its private slab of shared memory (demand-driven: the supply runs just ahead of
// the consumers); once the stage-1 queue is empty it waits.  Every incomplete plane it may wait for was claimed
// earlier by a warp that is running, and stage-1 tasks never wait, so there is no deadlock for any t_eval order.
// B-tiles are THIN (a few rows of g_2, only the rows of u they touch are staged) and the chunk LONG (32 outputs), so
// that the planes of T in use (~40 MB) stay in the 126 MB L2 between the stores of stage 1 and the loads of stage 2
// while the axis-4 walk keeps its length.  Completion is published per (tile, plane) with a device-scope counter
// (store T; __syncwarp; __threadfence + atomicAdd by one lane) and consumed by volatile loads + __threadfence; T is
// then read with ordinary (coherent) loads.
// MEASURED (B200, config 4, 128^4 outputs; two launches: 1.00 ms):
//   per-CTA scheduling, 8-output chunks, 2 fat B-tiles 1.64 ms; 32-output chunks, no tiling 1.17 ms;
//   per-CTA scheduling, thin B-tiles, 32-output chunks 1.51 ms (ncu profiles/r2_c4_fused.md: the T READ is gone from
//     DRAM — 0.87 GB read instead of 1.55 — but DRAM is 33 % busy and as many warps wait at __syncthreads as on memory);
//   per-warp scheduling (this version)             1.33 ms, the same for 20 ... 200 MB of live T.
// The saved traffic buys nothing because neither stage is DRAM-bound on its own (stage 2: issue 59 %, FP64 37 %,
// DRAM 55 %): the evaluation is bound by instruction issue and FP64 dependency latency at 16 warps/SM, and the
// persistent form adds queue / flag traffic and runs stage 1 at a worse shape.  The two-launch path stays the default.
template <typename T>
struct FusedParams {
    SliceParams<T> S;
    BatchParams<T> Q;        // Q.chunk / Q.n_bseg / Q.n_items are per B-tile
    int n_slices, n3;        // slices per tile, slices per plane
    int n_bt, rows_per_tile; // B-tiles = ranges of rows_per_tile outputs along g_2
    int s1_sub;              // g_2 outputs per lane and step of a stage-1 task
    int s1_piece;            // g_2 outputs of a tile staged together (their rows of u fit the warp's shared-memory slab)
    size_t warp_smem_elems;  // elements of T per warp in dynamic shared memory
    int n_chunks4, bx_per;   // stage-2 tasks = n_bt x n_chunks4 x bx_per
    int* ctrl;               // [0] stage-1 queue head, [1] stage-2 queue head, [2 + h * n4 + i_4] finished slices of (tile, plane)
};

constexpr size_t FUSED_CTRL_BYTES = 64 << 10;   // queue heads + per (tile, plane) completion counters, at the end of the scratch
constexpr int FU_BLOCK = S1_BLOCK;   // 256 threads: 8 stage-2 items per task
constexpr int FU_MAXCHUNK = 32;

template <typename T, int W, bool NURBS>
__global__ void __launch_bounds__(FU_BLOCK, 2) grid_split_fused_kernel(const __grid_constant__ FusedParams<T> F) {
    // Everything is WARP-granular: a warp claims a stage-2 item (64 batch columns x 2 rows of g_3 x one chunk of the
    // last axis), supplies missing planes by executing stage-1 tasks on its own (its private slab of shared memory),
    // and never meets a CTA barrier — the first version scheduled per CTA and ncu showed as many warps waiting at
    // __syncthreads (1.9 per issue cycle) as on memory.
    extern __shared__ __align__(16) unsigned char s1_smem[];
    __shared__ int s_i0[FU_BLOCK / 32][FU_MAXCHUNK];
    __shared__ T s_w[FU_BLOCK / 32][FU_MAXCHUNK * W];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* S = reinterpret_cast<T*>(s1_smem) + (size_t)warp * F.warp_smem_elems;
    constexpr int NARR = NURBS ? 2 : 1;
    const int n4 = F.Q.n4;
    const int per_tc = F.Q.n_items;                                   // stage-2 items per (tile, chunk)
    const int n_tasks2 = F.n_bt * F.n_chunks4 * per_tc, n_tasks1 = F.n_bt * F.n_slices;
    volatile int* done = F.ctrl + 2;
    for (;;) {
        int t2 = 0;
        if (lane == 0) t2 = atomicAdd(F.ctrl + 1, 1);
        t2 = __shfl_sync(0xffffffffu, t2, 0);
        if (t2 >= n_tasks2) break;
        const int h = t2 / (F.n_chunks4 * per_tc);
        const int r = t2 - h * (F.n_chunks4 * per_tc);
        const int c = r / per_tc, item = r - c * per_tc;
        for (;;) {
            // planes the chunk needs: [min, max + W) of its first nodes; lane l looks at output l and at plane l
            const int g_lo = c * F.Q.chunk, g_hi = min(F.Q.g4, g_lo + F.Q.chunk);
            const int p0 = __ldg(F.Q.i0_4 + min(g_lo + lane, g_hi - 1));
            const int p_lo = __reduce_min_sync(0xffffffffu, p0), p_hi = __reduce_max_sync(0xffffffffu, p0) + W;
            int ready = 1;
            for (int p = p_lo + lane; p < p_hi; p += 32) ready = ready && done[h * n4 + p] >= F.n3;
            ready = __all_sync(0xffffffffu, ready);
            if (ready) {
                __threadfence();
                break;
            }
            int s1 = 0;
            if (lane == 0) s1 = atomicAdd(F.ctrl, 1);
            s1 = __shfl_sync(0xffffffffu, s1, 0);
            if (s1 >= n_tasks1) {      // nothing left to supply: the missing planes are being written by other warps
                __nanosleep(200);
                continue;
            }
            const int h1 = s1 / F.n_slices, sl = s1 - h1 * F.n_slices;
            const int t_lo = h1 * F.rows_per_tile, t_hi = min(F.S.g2, t_lo + F.rows_per_tile);
            // the tile's outputs in pieces whose rows of u fit the warp's shared-memory slab
            for (int g2_lo = t_lo; g2_lo < t_hi; g2_lo += F.s1_piece) {
                const int g2_hi = min(t_hi, g2_lo + F.s1_piece);
                int lo = 1 << 30, hi = -1;
                for (int g = g2_lo + lane; g < g2_hi; g += 32) {
                    const int rr = __ldg(F.S.i0_2 + g);
                    lo = min(lo, rr);
                    hi = max(hi, rr);
                }
                lo = __reduce_min_sync(0xffffffffu, lo);
                hi = __reduce_max_sync(0xffffffffu, hi);
                int r_n = hi + W - lo;
                if ((size_t)r_n * F.S.n1 * NARR > F.warp_smem_elems) {   // rows too far apart (unsorted t_eval): one output at a time
                    for (int g = g2_lo; g < g2_hi; ++g) {
                        const int rr = __ldg(F.S.i0_2 + g);
                        slice_body<T, W, NARR, 32>(F.S, sl, g, g + 1, 1, rr, W, S, lane);
                        __syncwarp();
                    }
                } else {
                    slice_body<T, W, NARR, 32>(F.S, sl, g2_lo, g2_hi, F.s1_sub, lo, r_n, S, lane);
                    __syncwarp();
                }
            }
            if (lane == 0) {
                __threadfence();
                atomicAdd(F.ctrl + 2 + h1 * n4 + sl / F.n3, 1);
            }
        }
        batch_body<T, W, W, NURBS, 32, true>(F.Q, item, c, h * F.Q.n_bseg, s_i0[warp], s_w[warp], lane);
        __syncwarp();
    }
}

template <typename T, int W>
cudaError_t launch_slice(const SliceParams<T>& Q, int64_t n_slices, cudaStream_t s) {
    const int narr = Q.a1 ? 2 : 1;
    const size_t smem = (size_t)narr * Q.n1 * Q.n2 * sizeof(T);
    cudaError_t e;
    if (narr == 2) {
        if ((e = cudaFuncSetAttribute(grid_slice_kernel<T, W, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        grid_slice_kernel<T, W, 2><<<(unsigned)n_slices, S1_BLOCK, smem, s>>>(Q);
    } else {
        if ((e = cudaFuncSetAttribute(grid_slice_kernel<T, W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        grid_slice_kernel<T, W, 1><<<(unsigned)n_slices, S1_BLOCK, smem, s>>>(Q);
    }
    return cudaGetLastError();
}

template <typename T, int W, int W4>
cudaError_t launch_batch(const BatchParams<T>& Q0, const Knobs* knobs, cudaStream_t s) {
    BatchParams<T> Q = Q0;
    Q.n_bseg = (int)((Q.B + 32 * SB_VB - 1) / (32 * SB_VB));
    const int64_t items = (int64_t)Q.n_bseg * ((Q.g3 + 1) / 2);
    Q.n_items = (int)items;
    const int bx = (int)((items + SB_WARPS - 1) / SB_WARPS);
    int chunk = 32;
    while (chunk > 8 && (int64_t)bx * ((Q.g4 + chunk - 1) / chunk) < 148 * 4 * 2) chunk >>= 1;
    Q.chunk = chunk;
    Q.pf = (knobs && knobs->pf1 > 0) ? knobs->pf1 : SB_PF;
    dim3 grid((unsigned)bx, (unsigned)((Q.g4 + chunk - 1) / chunk));
    // adjacent column pairs (one vector load / store per pair) need an even B and pair-aligned arrays
    const size_t pair = 2 * sizeof(T);
    const bool vec = Q.B % 2 == 0 && (uintptr_t)Q.tn % pair == 0 && (uintptr_t)Q.td % pair == 0 && (uintptr_t)Q.out % pair == 0 &&
                     !(knobs && knobs->no_batch_vec);
    if (Q.td) {
        if (vec) grid_batch_kernel<T, W, W4, true, true><<<grid, SB_BLOCK, 0, s>>>(Q);
        else grid_batch_kernel<T, W, W4, true, false><<<grid, SB_BLOCK, 0, s>>>(Q);
    } else {
        if (vec) grid_batch_kernel<T, W, W4, false, true><<<grid, SB_BLOCK, 0, s>>>(Q);
        else grid_batch_kernel<T, W, W4, false, false><<<grid, SB_BLOCK, 0, s>>>(Q);
    }
    return cudaGetLastError();
}

// tiles / sizes of the fused launch; false = not applicable (then the two kernels run one after the other)
template <typename T>
bool fused_plan(const EvalParams<T>& P, FusedParams<T>& F, size_t ctrl_ints_cap) {
    const int W = P.dim[0].width;
    if (!(P.knobs && P.knobs->persist)) return false;   // opt-in (DIND_PERSIST=1): measured slower than the two launches, see above
    if (P.n_comp != 1 || P.dim[2].width != W || P.dim[3].width != W || !grid_split_slice_ok<T>(P)) return false;
    const int64_t g1 = P.dim[0].n_eval, g2 = P.dim[1].n_eval;
    const int n3 = P.dim[2].n_u, n4 = P.dim[3].n_u;
    const int narr = P.wts ? 2 : 1;
    // stage-2 chunk along the last axis: long (32 outputs), so that the window start-up of the walk is amortised;
    // B-tiles THIN (a few rows of g_2), so that the planes of T a chunk needs — about chunk * n4 / g4 + W of them —
    // stay within ~40 MB of L2 for the whole tile (experiment knobs: DIND_CHUNK, DIND_PF1 = MB)
    int chunk = (P.knobs && P.knobs->chunk > 0) ? P.knobs->chunk : 32;
    chunk = std::max(1, std::min(chunk, FU_MAXCHUNK));
    const double planes = (double)chunk * n4 / (double)P.dim[3].n_eval + W;
    const double budget = ((P.knobs && P.knobs->pf1 > 0) ? P.knobs->pf1 : 40) * 1.0e6;
    int rows = (int)g2;
    while (rows > 1 && ((double)g1 * rows * n3 * planes * sizeof(T) * narr > budget || g2 % rows != 0)) --rows;
    const int n_bt = (int)(g2 / rows);
    const int64_t b_tile = g1 * rows;
    if (b_tile % (32 * SB_VB) != 0) return false;
    if ((size_t)(2 + n_bt * n4) > ctrl_ints_cap) return false;
    F.n_slices = n3 * n4;
    F.n3 = n3;
    F.n_bt = n_bt;
    F.rows_per_tile = rows;
    // stage 1 runs per warp: its slab of shared memory (2 CTAs/SM x 8 warps share ~200 KB) holds the rows of u that a
    // piece of the tile's g_2 outputs touches — about piece * n2 / g2 + W rows of n1 elements (x2 for NURBS)
    const int n1 = P.dim[0].n_u, n2 = P.dim[1].n_u;
    F.warp_smem_elems = (size_t)(100 * 1024 / (FU_BLOCK / 32)) / sizeof(T);
    if ((size_t)W * n1 * narr > F.warp_smem_elems) return false;
    int piece = rows;
    while (piece > 1 && ((double)piece * n2 / (double)g2 + W + 1) * n1 * narr > (double)F.warp_smem_elems) piece = (piece + 1) / 2;
    F.s1_piece = piece;
    F.s1_sub = std::min(piece, 8);
    F.Q.chunk = chunk;
    F.Q.pf = SB_PF;
    F.n_chunks4 = (int)((P.dim[3].n_eval + chunk - 1) / chunk);
    F.Q.n_bseg = (int)(b_tile / (32 * SB_VB));
    const int64_t items = (int64_t)F.Q.n_bseg * ((P.dim[2].n_eval + 1) / 2);
    F.Q.n_items = (int)items;
    F.bx_per = (int)items;   // stage-2 tasks are single items (one warp each)
    return (int64_t)n_bt * F.n_chunks4 * items < (1LL << 30);
}

}  // namespace

template <typename T>
bool grid_split_supported(const EvalParams<T>& P) {
    if (P.n_in != 4) return false;
    const int W = P.dim[0].width;
    if (W != 3 && W != 4) return false;
    int64_t B = 1, items2 = 1;
    for (int d = 0; d < 4; ++d) {
        if (P.dim[d].n_eval < 1 || P.dim[d].n_eval >= 65535 || P.dim[d].n_u >= (1 << 20)) return false;
        // axes 1, 2: B-spline of one width; axes 3, 4: that width, or a Constant axis (width 1: a slice selector)
        if (P.dim[d].width != W && !(d >= 2 && P.dim[d].width == 1)) return false;
    }
    B = P.dim[0].n_eval * P.dim[1].n_eval;
    items2 = ((B + 63) / 64) * ((P.dim[2].n_eval + 1) / 2);
    if (B >= (1LL << 31) || items2 >= (1LL << 31) || P.u_comp_stride >= (1LL << 31)) return false;
    // worth it when the outputs outnumber the nodes along the first two axes (T is not larger than out)
    return P.dim[2].n_eval * P.dim[3].n_eval >= (int64_t)P.dim[2].n_u * P.dim[3].n_u;
}

// stage 1 can run from shared memory: the (n_1, n_2) slices (x2 for NURBS) fit and are contiguous in u
template <typename T>
bool grid_split_slice_ok(const EvalParams<T>& P) {
    const int64_t n1 = P.dim[0].n_u, n2 = P.dim[1].n_u, n3 = P.dim[2].n_u;
    const size_t slice_bytes = (size_t)n1 * n2 * sizeof(T) * (P.wts ? 2 : 1);
    return slice_bytes <= 96 * 1024 && P.dim[0].u_stride == 1 && P.dim[1].u_stride == n1 && P.dim[2].u_stride == n1 * n2 &&
           P.dim[3].u_stride == n1 * n2 * n3 && P.dim[0].n_eval < (1 << 24) && !(P.knobs && P.knobs->no_slice);
}

template <typename T>
size_t grid_split_scratch_bytes(const EvalParams<T>& P) {
    const size_t t_elems = (size_t)P.dim[0].n_eval * P.dim[1].n_eval * P.dim[2].n_u * P.dim[3].n_u;
    const size_t t_bytes = (t_elems * sizeof(T) + 255) & ~size_t(255);
    const size_t iota = ((size_t)std::max(P.dim[2].n_u, P.dim[3].n_u) * sizeof(int32_t) + 255) & ~size_t(255);
    return iota + t_bytes * (P.wts ? 2 : 1) + FUSED_CTRL_BYTES;
}

template <typename T>
cudaError_t launch_grid_split(const EvalParams<T>& P, void* scratch, cudaStream_t s, const char** name, int* launches) {
    const int W = P.dim[0].width;
    const int n3 = P.dim[2].n_u, n4 = P.dim[3].n_u;
    const int64_t g1 = P.dim[0].n_eval, g2 = P.dim[1].n_eval;
    const int64_t B = g1 * g2;
    const size_t t_elems = (size_t)B * n3 * n4;
    const size_t t_bytes = (t_elems * sizeof(T) + 255) & ~size_t(255);
    const int n_iota = std::max(n3, n4);
    const size_t iota_bytes = ((size_t)n_iota * sizeof(int32_t) + 255) & ~size_t(255);
    int32_t* iota = (int32_t*)scratch;
    T* tn = (T*)((char*)scratch + iota_bytes);
    T* td = P.wts ? (T*)((char*)scratch + iota_bytes + t_bytes) : nullptr;
    iota_kernel<<<(n_iota + 255) / 256, 256, 0, s>>>(iota, n_iota);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    *launches = 1;

    // stage 1.  Preferred: the shared-memory slice kernel (u*w and w in one pass); it needs the (n_1, n_2)
    // slice(s) to fit in shared memory and to be contiguous in u.  Otherwise: the 2-D pencil kernel over axes
    // (1, 2) with axes 3, 4 as identity "point" axes that select slices of u one to one.
    const int n1 = P.dim[0].n_u, n2 = P.dim[1].n_u;
    const bool use_slice = grid_split_slice_ok<T>(P);
    if (P.u_is_raw && !use_slice) return cudaErrorInvalidValue;
    EvalParams<T> P1 = P;
    P1.n_in = 4;
    P1.n_stencil = 2;
    P1.dim[0].out_stride = 1;
    P1.dim[1].out_stride = g1;
    for (int d = 2; d < 4; ++d) {
        P1.dim[d].width = 1;
        P1.dim[d].n_eval = d == 2 ? n3 : n4;
        P1.dim[d].tab_i0 = iota;
        P1.dim[d].tab_w = nullptr;
    }
    P1.dim[2].out_stride = B;
    P1.dim[3].out_stride = B * n3;
    P1.wts = nullptr;
    P1.n_comp = 1;
    P1.n_points = (int64_t)t_elems;
    P1.out_comp_stride = (int64_t)t_elems;
    const char* nm = nullptr;
    int nl = 0;
    SliceParams<T> S1;
    S1.n1 = n1; S1.n2 = n2; S1.g1 = (int)g1; S1.g2 = (int)g2;
    S1.slice_stride = (int64_t)n1 * n2;
    S1.i0_1 = P.dim[0].tab_i0; S1.w_1 = P.dim[0].tab_w;
    S1.i0_2 = P.dim[1].tab_i0; S1.w_2 = P.dim[1].tab_w;
    // both stages in one persistent launch (scalar u, all four axes B-splines, slices fit shared memory)
    {
        FusedParams<T> F;
        memset(&F, 0, sizeof F);
        F.S = S1;
        F.S.a0 = P.u;
        F.S.t0 = tn;
        F.S.a1 = td ? P.wts : nullptr;
        F.S.premul = P.u_is_raw ? P.wts : nullptr;
        F.S.t1 = td;
        F.Q.tn = tn; F.Q.td = td; F.Q.out = P.out; F.Q.B = B; F.Q.n3 = n3; F.Q.n4 = n4;
        F.Q.g3 = (int)P.dim[2].n_eval; F.Q.g4 = (int)P.dim[3].n_eval;
        F.Q.i0_3 = P.dim[2].tab_i0; F.Q.w_3 = P.dim[2].tab_w; F.Q.i0_4 = P.dim[3].tab_i0; F.Q.w_4 = P.dim[3].tab_w;
        F.ctrl = (int*)((char*)scratch + iota_bytes + t_bytes * (td ? 2 : 1));
        if (fused_plan<T>(P, F, FUSED_CTRL_BYTES / sizeof(int))) {
            const size_t ctrl_bytes = (size_t)(2 + F.n_bt * n4) * sizeof(int);
            if ((e = cudaMemsetAsync(F.ctrl, 0, ctrl_bytes, s)) != cudaSuccess) return e;
            const size_t smem = F.warp_smem_elems * sizeof(T) * (FU_BLOCK / 32);
            int dev = 0, n_sm = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            const int grid = 2 * n_sm;
#define DIND_FUSED(WW, NU)                                                                                                  \
    {                                                                                                                      \
        if ((e = cudaFuncSetAttribute(grid_split_fused_kernel<T, WW, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                      (int)smem)) != cudaSuccess) return e;                                                \
        grid_split_fused_kernel<T, WW, NU><<<grid, FU_BLOCK, smem, s>>>(F);                                                 \
    }
            if (W == 3 && td) DIND_FUSED(3, true)
            else if (W == 3) DIND_FUSED(3, false)
            else if (td) DIND_FUSED(4, true)
            else DIND_FUSED(4, false)
#undef DIND_FUSED
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            *launches += 1;
            *name = W == 3 ? "grid_split_fused<N=2+2,W=3>" : "grid_split_fused<N=2+2,W=4>";
            return cudaSuccess;
        }
    }
    bool den_done = td && P.split_den_cached;   // the weights' stage-1 result of an earlier evaluation is still valid
    if (td && !use_slice && !den_done) {
        P1.u = P.wts;
        P1.out = td;
        if (!try_eval_fast_grid<T>(P1, s, &nm, &e, &nl)) return cudaErrorNotSupported;
        if (e != cudaSuccess) return e;
        *launches += nl;
        den_done = true;
    }
    for (int c = 0; c < P.n_comp; ++c) {
        if (use_slice) {
            S1.a0 = P.u + (int64_t)c * P.u_comp_stride;
            S1.t0 = tn;
            S1.a1 = (td && !den_done) ? P.wts : nullptr;   // the denominator rides along with the first component
            S1.premul = P.u_is_raw ? P.wts : nullptr;
            S1.t1 = td;
            e = W == 3 ? launch_slice<T, 3>(S1, (int64_t)n3 * n4, s) : launch_slice<T, 4>(S1, (int64_t)n3 * n4, s);
            if (e != cudaSuccess) return e;
            den_done = true;
            *launches += 1;
        } else {
            P1.u = P.u + (int64_t)c * P.u_comp_stride;
            P1.out = tn;
            if (!try_eval_fast_grid<T>(P1, s, &nm, &e, &nl)) return cudaErrorNotSupported;
            if (e != cudaSuccess) return e;
            *launches += nl;
        }
        // stage 2
        BatchParams<T> Q;
        Q.tn = tn;
        Q.td = td;
        Q.out = P.out + (int64_t)c * P.out_comp_stride;
        Q.B = B;
        Q.n3 = n3;
        Q.n4 = n4;
        Q.g3 = (int)P.dim[2].n_eval;
        Q.g4 = (int)P.dim[3].n_eval;
        Q.i0_3 = P.dim[2].tab_i0;
        Q.w_3 = P.dim[2].tab_w;
        Q.i0_4 = P.dim[3].tab_i0;
        Q.w_4 = P.dim[3].tab_w;
        Q.chunk = 0;
        Q.n_bseg = Q.n_items = 0;
        const int W3 = P.dim[2].width, W4 = P.dim[3].width;
#define DIND_BATCH(A, B4) \
    if (W3 == A && W4 == B4) e = launch_batch<T, A, B4>(Q, P.knobs, s); else
        DIND_BATCH(3, 3) DIND_BATCH(4, 4) DIND_BATCH(1, 3) DIND_BATCH(3, 1) DIND_BATCH(1, 4) DIND_BATCH(4, 1) DIND_BATCH(1, 1)
        e = cudaErrorNotSupported;
#undef DIND_BATCH
        if (e != cudaSuccess) return e;
        *launches += 1;
    }
    *name = (P.dim[2].width == W && P.dim[3].width == W) ? (W == 3 ? "grid_split<N=2+2,W=3>" : "grid_split<N=2+2,W=4>")
                                                         : (W == 3 ? "grid_split<N=2+2,W=3,selector axes>" : "grid_split<N=2+2,W=4,selector axes>");
    return cudaSuccess;
}

#define DIND_INST(T)                                                                                       \
    template bool grid_split_supported<T>(const EvalParams<T>&);                                           \
    template bool grid_split_slice_ok<T>(const EvalParams<T>&);                                            \
    template size_t grid_split_scratch_bytes<T>(const EvalParams<T>&);                                     \
    template cudaError_t launch_grid_split<T>(const EvalParams<T>&, void*, cudaStream_t, const char**, int*);
DIND_INST(float)
DIND_INST(double)

}  // namespace dind
