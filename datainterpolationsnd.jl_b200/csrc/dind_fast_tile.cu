// dind_fast_tile.cu — eval_unstructured! over the cell-sorted plan's WORK ITEMS: one thread evaluates up to PT
// consecutive plan points that lie in the same cell and loads every stencil node ONCE for all of them.
//
// Why: in plan order every lane of a warp reads the same (degree+1)^N stencil, yet with one point per thread
// every lane still pulls its own copy through the L1 -> register path: BASELINE config 3 (3-D cubic B-spline,
// 3 components, Float64) moves 64 nodes x 24 B = 1.5 KB per point, 49 KB per warp of points = 384 clk of the
// 128 B/clk L1 data path against 189 clk of FP64 issue (ncu, profiles/r1_c3_sorted.md: l1tex 94 % busy, FP64
// pipe 33 %).  The only way to cut L1 -> register bytes per point is to reuse a loaded node in registers, i.e.
// several points of the SAME cell per thread.  A fixed "PT consecutive points per thread" split straddles cell
// boundaries (with ~51 points per cell and PT = 4, most warps would contain a straddling thread), so the plan
// carries a schedule instead (dind_plan.cu: plan_build_items): the plan positions are cut into items of at most
// PT points that never cross a cell boundary.  A thread takes item i = [items[i], items[i + 1]); slots beyond
// the item's length repeat its last point and are not stored.
//
// The per-point arithmetic (weights, FMA order of the axis-by-axis contraction) is exactly that of
// unstructured_aos_kernel, so results are bit-identical to the unsorted evaluation.
//
// Reference: eval_unstructured! / eval_kernel (src/interpolation_parallel.jl:34-52, :105-138) evaluate one point
// per work item and redo the whole stencil gather for each.
#include <type_traits>

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int TBLOCK = 128;

// same per-dimension stencil as unstructured_aos_kernel (dind_fast_unstructured_aos.cu)
template <typename T, int W>
__device__ __forceinline__ void tile_weights(const DimParams<T>& D, T x, int idx, int order, T (&w)[W]) {
    const int n = D.n_knots;
    if (W == 2 && D.kind == LINEAR) {
        const int c = idx - 1;
        const T t1 = __ldg(D.knots + c), t2 = __ldg(D.knots + c + 1), r = __ldg(D.recip + c);
        if (order == 0) { w[0] = (t2 - x) * r; w[1] = (x - t1) * r; }
        else { w[0] = -r; w[1] = r; }
        return;
    }
    bspline_basis_ct<T, W - 1>(D.knots, D.recip, n, idx - 1, x, order, w);
}

// Axis-by-axis contraction for PT points that share the stencil: node loads are shared, the FMA chains are per point
// (and per point identical to ContractV of dind_fast_unstructured_aos.cu).
// SMEM: the stencil was staged in shared memory (unstructured_tile_smem_kernel): plain loads instead of the read-only path
template <typename T, int C>
__device__ __forceinline__ Vec<T, C> load_node_smem(const T* p) {
    Vec<T, C> r;
    if constexpr (sizeof(T) == 8 && C >= 2) {
        const double2 a = *reinterpret_cast<const double2*>(p);
        r.v[0] = a.x;
        r.v[1] = a.y;
        if constexpr (C == 3) r.v[2] = p[2];
        if constexpr (C == 4) {
            const double2 b = *reinterpret_cast<const double2*>(p + 2);
            r.v[2] = b.x;
            r.v[3] = b.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) r.v[c] = p[c];
    }
    return r;
}

template <typename T, int D, int W, int C, int PT, bool SMEM = false>
struct ContractT {
    static __device__ __forceinline__ void run(const T* __restrict__ p, const T (&w)[PT][MAXN][W], const int (&stride)[MAXN],
                                               T (&acc)[PT][C]) {
#pragma unroll
        for (int q = 0; q < PT; ++q)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[q][c] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            T t[PT][C];
            ContractT<T, D - 1, W, C, PT, SMEM>::run(p + a * stride[D], w, stride, t);
#pragma unroll
            for (int q = 0; q < PT; ++q)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[q][c] = fma(w[q][D][a], t[q][c], acc[q][c]);
        }
    }
};
template <typename T, int W, int C, int PT, bool SMEM>
struct ContractT<T, 0, W, C, PT, SMEM> {
    static __device__ __forceinline__ void run(const T* __restrict__ p, const T (&w)[PT][MAXN][W], const int (&)[MAXN],
                                               T (&acc)[PT][C]) {
        constexpr int S = C == 3 ? 4 : C;
        Vec<T, C> node[W];
#pragma unroll
        for (int a = 0; a < W; ++a) {   // all loads of the row first
            if constexpr (SMEM) node[a] = load_node_smem<T, C>(p + a * S);
            else node[a] = load_node<T, C, false>(p + a * S);
        }
#pragma unroll
        for (int q = 0; q < PT; ++q)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[q][c] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a)
#pragma unroll
            for (int q = 0; q < PT; ++q)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[q][c] = fma(w[q][0][a], node[a].v[c], acc[q][c]);
    }
};

// Loads the compiler must issue where they are written (volatile asm keeps its place among the other volatile
// statements): left alone, ptxas sinks the knot / output-slot loads to their first use to save registers, and every
// one of them then costs a full dependent round trip (ncu source view of linear_tile_kernel: 13 % of all stall
// samples on the first FADD of the last dimension's weights, 5 % on the store address).
__device__ __forceinline__ float ld_pinned(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_pinned(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_pinned(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void tile_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tile_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Inputs are ITEM-MAJOR (dind_plan.cu: plan_item_layout): P.cells[i] is item i's cell word and P.dim[d].t_eval[i * PT + q]
// the coordinate of its q-th point (padding slots repeat the last point), so every input of a thread is addressed by
// the thread index alone — one DRAM round trip instead of items -> points — and the inputs of the CTA `pf_ctas`
// further on can be requested into L2 now.  P.items (item i = plan positions [items[i], items[i + 1])) is only needed
// for the output position.
// DER: the derivative orders known at compile time — -1: all zero; d >= 0: order 1 in dimension d, 0 elsewhere (the four
// evaluations of "value plus first derivatives"); -2: read P.dim[d].order at run time.  With the order a constant the
// value / derivative branches of every Cox-de Boor stage fold away instead of being predicated (ncu: a third of the
// FP64 instructions of the basis were predicated off).
template <typename T, int N, int W, int C, int PT, int MINB, int DER>
__global__ void __launch_bounds__(TBLOCK, MINB) unstructured_tile_kernel(const __grid_constant__ EvalParams<T> P, int pf_ctas) {
    const int64_t i = (int64_t)blockIdx.x * TBLOCK + threadIdx.x;
    if (i >= P.n_items) return;
    constexpr int S = C == 3 ? 4 : C;
    const uint32_t cell = __ldcs(P.cells + i);
    T xs[PT][N];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q) xs[q][d] = __ldcs(P.dim[d].t_eval + i * PT + q);
    const uint32_t s0 = __ldcs(P.items + i), s1 = __ldcs(P.items + i + 1);
    {   // L2 prefetch of the same inputs for the CTA pf_ctas ahead (contiguous per warp: a few lines per instruction)
        const int64_t j = i + (int64_t)pf_ctas * TBLOCK;
        if (pf_ctas > 0 && j < P.n_items) {
            tile_prefetch_l2(P.cells + j);
            tile_prefetch_l2(P.items + j);
#pragma unroll
            for (int d = 0; d < N; ++d)
#pragma unroll
                for (int q = 0; q < PT; ++q) tile_prefetch_l2(P.dim[d].t_eval + j * PT + q);
        }
    }
    int idx[N];
    int stride[MAXN];
    int64_t base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        idx[d] = cell_idx<T>(D, cell);
        stride[d] = (int)D.u_stride * S;
        base += (int64_t)(idx[d] - (W == 2 && D.kind == LINEAR ? 1 : W)) * D.u_stride;
    }
    const T* ub = P.u_aos + base * S;
    if constexpr (N == 3 && W == 4) {
        // the stencil's 16 rows of 4 nodes (128 contiguous bytes each, usually across two lines) into L1 while the
        // weights are computed: a group of 4 neighbouring lanes (same cell, nearly always) covers all 32 (row, half)
        const int lane = threadIdx.x & 31;
        const int half = lane & 1, r0 = ((lane >> 1) & 1) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = r0 + j;
            tile_prefetch_l1(ub + (r & 3) * stride[1] + (r >> 2) * stride[2] + half * 3 * S);
        }
    }
    T w[PT][MAXN][W];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q)
            tile_weights<T, W>(P.dim[d], xs[q][d], idx[d], DER == -2 ? P.dim[d].order : (DER == d ? 1 : 0), w[q][d]);
    const int cnt = (int)(s1 - s0);                     // 1..PT points, all in one cell
    uint32_t ko[PT];
#pragma unroll
    for (int q = 0; q < PT; ++q) ko[q] = (P.perm && q < cnt) ? __ldcs(P.perm + s0 + q) : s0 + q;
    T acc[PT][C];
    ContractT<T, N - 1, W, C, PT>::run(ub, w, stride, acc);
#pragma unroll
    for (int q = 0; q < PT; ++q) {
        if (q >= cnt) break;
        if (P.staged_records) {   // one aligned record per point (the plan's staging buffer)
            store_record<T, C>(P.out + (int64_t)ko[q] * P.out_point_stride, acc[q]);
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) __stcs(P.out + (int64_t)c * P.out_comp_stride + (int64_t)ko[q] * P.out_point_stride, acc[q][c]);
        }
    }
}

// ---- the same kernel with the stencils staged in shared memory ---------------------------------------------------
// ncu on the kernel above (config 3 in plan order, profiles/r2_c3_plan_order.md): FP64 pipe 56 % busy, 39 % of the L1
// requests miss and the warps (3 per scheduler at 168 registers) wait for L2 on the nodes their FMA chains need —
// long_scoreboard is the largest stall.  The items of a warp are consecutive in the plan, so they belong to two or
// three consecutive cells (17 items per cell at 1e8 points).  Here every WARP finds its distinct cells (flag = the
// cell differs from the previous lane's; ballot), copies each cell's 4 x 4 x 4 stencil (16 rows of 128 contiguous
// bytes) from the component-fastest copy of u into a warp-private shared-memory slot with cp.async — issued before
// the basis functions are computed, waited for after — and the contraction reads nodes from shared memory (lanes of
// one cell read one address: broadcast; slots are 32 bytes apart modulo the bank cycle, so the cells of a warp do
// not conflict).  No CTA barrier: warps stay decoupled and the FMA phase never waits on L2.  Same weights, same FMA
// chains: bit-identical.  More distinct cells in a warp than slots (sparse plans): rounds of TS_SLOTS slots; the
// launcher only picks this kernel where that is rare.  Nodes of 32 bytes only (Float64, 3 or 4 components).
constexpr int TS_SLOTS = 6;   // per warp
template <typename T, int C> struct TileSmem {
    static constexpr int S = C == 3 ? 4 : C;
    static constexpr int SLOT_ELEMS = 64 * S + 32 / (int)sizeof(T);   // 2048 + 32 bytes
    static constexpr size_t BYTES = (size_t)(TBLOCK / 32) * TS_SLOTS * SLOT_ELEMS * sizeof(T);
};

// The warp's distinct cells and their shared-memory slots (see above).
template <typename T, int C>
struct WarpStencils {
    static constexpr int S = TileSmem<T, C>::S, SLOT_ELEMS = TileSmem<T, C>::SLOT_ELEMS;
    T* sm;            // the warp's TS_SLOTS slots
    unsigned rem;     // heads whose cells have not been staged yet
    int slot, total;  // this lane's slot; distinct cells of the warp
    int64_t ch_src[4];
    unsigned ch_dst[4];
    __device__ __forceinline__ void init(unsigned char* smem, uint32_t cell, int64_t st1, int64_t st2) {
        const int lane = threadIdx.x & 31;
        sm = reinterpret_cast<T*>(smem) + (threadIdx.x >> 5) * (TS_SLOTS * SLOT_ELEMS);
        const uint32_t prev = __shfl_up_sync(0xffffffffu, cell, 1);
        const bool head = lane == 0 || cell != prev;
        rem = __ballot_sync(0xffffffffu, head);
        slot = __popc(rem & (0xffffffffu >> (31 - lane))) - 1;
        total = __popc(rem);
        // a stencil = 16 rows x 128 bytes = 128 chunks of 16 bytes: four per lane
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ch = lane + 32 * j, row = ch >> 3, part = (ch & 7) * (16 / (int)sizeof(T));
            ch_src[j] = ((row & 3) * st1 + (row >> 2) * st2) * S + part;
            ch_dst[j] = (unsigned)__cvta_generic_to_shared(sm + row * (4 * S) + part);
        }
    }
    // cp.async of the cells [first, first + TS_SLOTS) of the warp; base = this lane's first stencil node
    __device__ __forceinline__ void stage(const T* __restrict__ u_aos, int64_t base, int first) {
        const int last = min(total, first + TS_SLOTS);
        for (int sl = first; sl < last; ++sl) {
            const int64_t b = __shfl_sync(0xffffffffu, base, __ffs(rem) - 1);   // (__fns is a software loop: 32.4 vs 34.5 Gpoints/s)
            rem &= rem - 1;
            const T* src = u_aos + b * S;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ch_dst[j] + (unsigned)((sl - first) * SLOT_ELEMS * sizeof(T))),
                             "l"(src + ch_src[j])
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void wait() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ bool mine(int first) const { return slot >= first && slot < first + TS_SLOTS; }
    __device__ __forceinline__ const T* stencil(int first) const { return sm + (slot - first) * SLOT_ELEMS; }
};

template <typename T, int C, int PT, int MINB, int DER>
__global__ void __launch_bounds__(TBLOCK, MINB) unstructured_tile_smem_kernel(const __grid_constant__ EvalParams<T> P, int pf_ctas) {
    constexpr int N = 3, W = 4, S = TileSmem<T, C>::S;
    static_assert(S * sizeof(T) == 32, "32-byte nodes");
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int64_t i_raw = (int64_t)blockIdx.x * TBLOCK + threadIdx.x;
    const bool valid = i_raw < P.n_items;
    const int64_t i = valid ? i_raw : P.n_items - 1;     // threads past the end repeat the last item (no new cell, no store)
    const uint32_t cell = __ldcs(P.cells + i);
    T xs[PT][N];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q) xs[q][d] = __ldcs(P.dim[d].t_eval + i * PT + q);
    const uint32_t s0 = __ldcs(P.items + i), s1 = __ldcs(P.items + i + 1);
    {
        const int64_t j = i_raw + (int64_t)pf_ctas * TBLOCK;
        if (pf_ctas > 0 && j < P.n_items) {
            tile_prefetch_l2(P.cells + j);
            tile_prefetch_l2(P.items + j);
#pragma unroll
            for (int d = 0; d < N; ++d)
#pragma unroll
                for (int q = 0; q < PT; ++q) tile_prefetch_l2(P.dim[d].t_eval + j * PT + q);
        }
    }
    int idx[N];
    int64_t base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        idx[d] = cell_idx<T>(P.dim[d], cell);
        base += (int64_t)(idx[d] - W) * P.dim[d].u_stride;
    }
    WarpStencils<T, C> ws;
    ws.init(tile_smem, cell, P.dim[1].u_stride, P.dim[2].u_stride);
    ws.stage(P.u_aos, base, 0);
    T w[PT][MAXN][W];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q)
            tile_weights<T, W>(P.dim[d], xs[q][d], idx[d], DER == -2 ? P.dim[d].order : (DER == d ? 1 : 0), w[q][d]);
    const int cnt = valid ? (int)(s1 - s0) : 0;
    uint32_t ko[PT];
#pragma unroll
    for (int q = 0; q < PT; ++q) ko[q] = (P.perm && q < cnt) ? __ldcs(P.perm + s0 + q) : s0 + q;
    const int sstride[MAXN] = {S, 4 * S, 16 * S, 0, 0, 0, 0, 0};
    for (int first = 0; first < ws.total; first += TS_SLOTS) {
        ws.wait();
        if (ws.mine(first)) {
            T acc[PT][C];
            ContractT<T, N - 1, W, C, PT, true>::run(ws.stencil(first), w, sstride, acc);
#pragma unroll
            for (int q = 0; q < PT; ++q) {
                if (q >= cnt) break;
                if (P.staged_records) {
                    store_record<T, C>(P.out + (int64_t)ko[q] * P.out_point_stride, acc[q]);
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) __stcs(P.out + (int64_t)c * P.out_comp_stride + (int64_t)ko[q] * P.out_point_stride, acc[q][c]);
                }
            }
        }
        if (first + TS_SLOTS < ws.total) {   // sparse plan: another round of slots
            __syncwarp();
            ws.stage(P.u_aos, base, first + TS_SLOTS);
        }
    }
}

// ---- fused value + first partials over work items (dind_eval_unstructured_gradient) -----------------------------
// Same contraction as ContractGV of dind_fast_gradient.cu (per point the same FMAs in the same order: bit-identical
// results), node loads shared by the item's PT points.
template <typename T, int D, int W, int C, int PT, bool SMEM = false>
struct ContractGT {
    static __device__ __forceinline__ void run(const T* __restrict__ p, const T (&w)[PT][MAXN][W], const T (&dw)[PT][MAXN][W],
                                               const int (&stride)[MAXN], T (&r)[PT][(D + 2) * C]) {
#pragma unroll
        for (int q = 0; q < PT; ++q)
#pragma unroll
            for (int j = 0; j < (D + 2) * C; ++j) r[q][j] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a) {
            T sub[PT][(D + 1) * C];
            ContractGT<T, D - 1, W, C, PT, SMEM>::run(p + a * stride[D], w, dw, stride, sub);
#pragma unroll
            for (int q = 0; q < PT; ++q) {
#pragma unroll
                for (int j = 0; j < (D + 1) * C; ++j) r[q][j] = fma(w[q][D][a], sub[q][j], r[q][j]);
#pragma unroll
                for (int c = 0; c < C; ++c) r[q][(D + 1) * C + c] = fma(dw[q][D][a], sub[q][c], r[q][(D + 1) * C + c]);
            }
        }
    }
};
template <typename T, int W, int C, int PT, bool SMEM>
struct ContractGT<T, 0, W, C, PT, SMEM> {
    static __device__ __forceinline__ void run(const T* __restrict__ p, const T (&w)[PT][MAXN][W], const T (&dw)[PT][MAXN][W],
                                               const int (&)[MAXN], T (&r)[PT][2 * C]) {
        constexpr int S = C == 3 ? 4 : C;
        Vec<T, C> node[W];
#pragma unroll
        for (int a = 0; a < W; ++a) {
            if constexpr (SMEM) node[a] = load_node_smem<T, C>(p + a * S);
            else node[a] = load_node<T, C, false>(p + a * S);
        }
#pragma unroll
        for (int q = 0; q < PT; ++q)
#pragma unroll
            for (int j = 0; j < 2 * C; ++j) r[q][j] = T(0);
#pragma unroll
        for (int a = 0; a < W; ++a)
#pragma unroll
            for (int q = 0; q < PT; ++q)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    r[q][c] = fma(w[q][0][a], node[a].v[c], r[q][c]);
                    r[q][C + c] = fma(dw[q][0][a], node[a].v[c], r[q][C + c]);
                }
    }
};

// SM: stencils staged in warp-private shared memory (WarpStencils above; N = 3, W = 4, 32-byte nodes)
template <typename T, int N, int W, int C, int PT, int MINB, bool SM>
__global__ void __launch_bounds__(TBLOCK, MINB) gradient_tile_kernel(const __grid_constant__ EvalParams<T> P, int pf_ctas) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    const int64_t i_raw = (int64_t)blockIdx.x * TBLOCK + threadIdx.x;
    const bool valid = i_raw < P.n_items;
    if (!SM && !valid) return;
    const int64_t i = valid ? i_raw : P.n_items - 1;     // SM: whole warps stay (threads past the end repeat the last item, store nothing)
    constexpr int S = C == 3 ? 4 : C;
    const uint32_t cell = __ldcs(P.cells + i);
    T xs[PT][N];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q) xs[q][d] = __ldcs(P.dim[d].t_eval + i * PT + q);
    const uint32_t s0 = __ldcs(P.items + i), s1 = __ldcs(P.items + i + 1);
    {
        const int64_t j = i_raw + (int64_t)pf_ctas * TBLOCK;
        if (pf_ctas > 0 && j < P.n_items) {
            tile_prefetch_l2(P.cells + j);
            tile_prefetch_l2(P.items + j);
#pragma unroll
            for (int d = 0; d < N; ++d)
#pragma unroll
                for (int q = 0; q < PT; ++q) tile_prefetch_l2(P.dim[d].t_eval + j * PT + q);
        }
    }
    int idx[N];
    int stride[MAXN];
    int64_t base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        idx[d] = cell_idx<T>(D, cell);
        stride[d] = (int)D.u_stride * S;
        base += (int64_t)(idx[d] - (W == 2 && D.kind == LINEAR ? 1 : W)) * D.u_stride;
    }
    WarpStencils<T, C> ws;
    if constexpr (SM) {
        ws.init(tile_smem, cell, P.dim[1].u_stride, P.dim[2].u_stride);
        ws.stage(P.u_aos, base, 0);
    }
    T w[PT][MAXN][W], dw[PT][MAXN][W];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            tile_weights<T, W>(P.dim[d], xs[q][d], idx[d], 0, w[q][d]);
            tile_weights<T, W>(P.dim[d], xs[q][d], idx[d], 1, dw[q][d]);
        }
    const int cnt = valid ? (int)(s1 - s0) : 0;
    uint32_t ko[PT];
#pragma unroll
    for (int q = 0; q < PT; ++q) ko[q] = (P.perm && q < cnt) ? __ldcs(P.perm + s0 + q) : s0 + q;
    const int64_t slot = P.out_comp_stride * C;
    auto finish = [&](const T* stencil, const int (&st)[MAXN], auto smem_tag) {
        T g[PT][(N + 1) * C];
        ContractGT<T, N - 1, W, C, PT, decltype(smem_tag)::value>::run(stencil, w, dw, st, g);
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            if (q >= cnt) break;
            T* o = P.out + (int64_t)ko[q] * P.out_point_stride;
            if (P.staged_records) {
                store_record<T, (N + 1) * C>(o, g[q]);
            } else {
#pragma unroll
                for (int j = 0; j < N + 1; ++j)
#pragma unroll
                    for (int c = 0; c < C; ++c) __stcs(o + (int64_t)c * P.out_comp_stride + j * slot, g[q][j * C + c]);
            }
        }
    };
    if constexpr (SM) {
        const int sstride[MAXN] = {S, 4 * S, 16 * S, 0, 0, 0, 0, 0};
        for (int first = 0; first < ws.total; first += TS_SLOTS) {
            ws.wait();
            if (ws.mine(first)) finish(ws.stencil(first), sstride, std::true_type{});
            if (first + TS_SLOTS < ws.total) {
                __syncwarp();
                ws.stage(P.u_aos, base, first + TS_SLOTS);
            }
        }
    } else {
        finish(P.u_aos + base * S, stride, std::false_type{});
    }
}

// plan dense enough for the shared-memory variants: >= 8 items per cell on average (<= 5 cells per warp, TS_SLOTS = 6).
// Measured (config 3 shape, plan order, shared memory vs global): 2e7 points (3.4 items per cell) 17.7 vs 24.8 Gpoints/s,
// 5e7 (8.5) 29.4 vs 28.7, 1e8 (17) 34.5 vs 29.4.
template <typename T, int N, int W>
bool tile_smem_wanted(const EvalParams<T>& P) {
    double cells_total = 1.0;
    for (int d = 0; d < N; ++d) cells_total *= (double)(P.dim[d].n_u - W + 1);
    const int mode = P.knobs ? P.knobs->tile_smem : -1;
    return mode == 1 || (mode != 0 && (double)P.n_items >= 8.0 * cells_total);
}

template <typename T, int N, int W, int C, int PT, int MINB>
cudaError_t launch_gradient_tile(const EvalParams<T>& P, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_items + TBLOCK - 1) / TBLOCK);
    const int pf = (P.knobs && P.knobs->pf1 >= 0) ? P.knobs->pf1 : 2 * 148 * MINB;
    if constexpr (N == 3 && W == 4 && sizeof(T) == 8 && (C == 3 || C == 4)) {
        if (tile_smem_wanted<T, N, W>(P)) {
            const size_t smem = TileSmem<T, C>::BYTES;
            cudaError_t e = cudaFuncSetAttribute(gradient_tile_kernel<T, N, W, C, PT, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            gradient_tile_kernel<T, N, W, C, PT, MINB, true><<<nb, TBLOCK, smem, s>>>(P, pf);
            return cudaGetLastError();
        }
    }
    gradient_tile_kernel<T, N, W, C, PT, MINB, false><<<nb, TBLOCK, 0, s>>>(P, pf);
    return cudaGetLastError();
}

// ---- multilinear tables (BASELINE config 5: 5-D linear, 4 Float32 components) ------------------------------------
// The one-point-per-thread kernel is issue-bound in plan order (ncu, profiles/r1_c5_sorted.md: issue slots 59 % busy,
// 645 instructions per point of which 248 are FFMA and ~370 control flow / parameter loads / address arithmetic of
// the general kernel).  This kernel is specialised for it: Linear dimensions only (no kind / search branches), the
// item's points share the knot loads, the corner loads and the address arithmetic, and the four components of a
// node are contracted as two packed pairs — fma.rn.f32x2 (FFMA2), one issue slot for two FMAs.  Packed FMAs round each
// half exactly like the scalar fmaf of unstructured_aos_kernel and the operation order per point is the same, so the
// results are bit-identical to the unsorted evaluation.
template <int D, int PT>
struct LerpT {
    static __device__ __forceinline__ void run(const float4* __restrict__ p, const float (&w)[PT][MAXN][2], const int (&stride)[MAXN],
                                               float2 (&lo)[PT], float2 (&hi)[PT]) {
        float2 tl[PT], th[PT];
        LerpT<D - 1, PT>::run(p, w, stride, tl, th);
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const float2 ww = make_float2(w[q][D][0], w[q][D][0]);
            lo[q] = __ffma2_rn(ww, tl[q], make_float2(0.f, 0.f));
            hi[q] = __ffma2_rn(ww, th[q], make_float2(0.f, 0.f));
        }
        LerpT<D - 1, PT>::run(p + stride[D], w, stride, tl, th);
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const float2 ww = make_float2(w[q][D][1], w[q][D][1]);
            lo[q] = __ffma2_rn(ww, tl[q], lo[q]);
            hi[q] = __ffma2_rn(ww, th[q], hi[q]);
        }
    }
};
template <int PT>
struct LerpT<0, PT> {
    static __device__ __forceinline__ void run(const float4* __restrict__ p, const float (&w)[PT][MAXN][2], const int (&)[MAXN],
                                               float2 (&lo)[PT], float2 (&hi)[PT]) {
        const float4 n0 = __ldg(p), n1 = __ldg(p + 1);
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const float2 w0 = make_float2(w[q][0][0], w[q][0][0]), w1 = make_float2(w[q][0][1], w[q][0][1]);
            lo[q] = __ffma2_rn(w1, make_float2(n1.x, n1.y), __ffma2_rn(w0, make_float2(n0.x, n0.y), make_float2(0.f, 0.f)));
            hi[q] = __ffma2_rn(w1, make_float2(n1.z, n1.w), __ffma2_rn(w0, make_float2(n0.z, n0.w), make_float2(0.f, 0.f)));
        }
    }
};

template <int N, int PT, int MINB, int DER>
__global__ void __launch_bounds__(TBLOCK, MINB) linear_tile_kernel(const __grid_constant__ EvalParams<float> P, int pf_ctas) {
    const int64_t i = (int64_t)blockIdx.x * TBLOCK + threadIdx.x;
    if (i >= P.n_items) return;
    const uint32_t cell = ld_pinned(P.cells + i);
    float xs[PT][N];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int q = 0; q < PT; ++q) xs[q][d] = ld_pinned(P.dim[d].t_eval + i * PT + q);
    const uint32_t s0 = ld_pinned(P.items + i), s1 = ld_pinned(P.items + i + 1);
    {
        const int64_t j = i + (int64_t)pf_ctas * TBLOCK;
        if (pf_ctas > 0 && j < P.n_items) {
            tile_prefetch_l2(P.cells + j);
            tile_prefetch_l2(P.items + j);
#pragma unroll
            for (int d = 0; d < N; ++d)
#pragma unroll
                for (int q = 0; q < PT; ++q) tile_prefetch_l2(P.dim[d].t_eval + j * PT + q);
        }
    }
    float w[PT][MAXN][2];
    int stride[MAXN];
    int64_t base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        stride[d] = (int)P.dim[d].u_stride;                 // in nodes = float4 units
        base += (int64_t)(cell_idx<float>(P.dim[d], cell) - 1) * P.dim[d].u_stride;
    }
    const float4* ub = reinterpret_cast<const float4*>(P.u_aos) + base;
    float kt1[N], kt2[N], kr[N];   // knots and reciprocal span of the item's cell: once per item, all requested now
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int c = cell_idx<float>(P.dim[d], cell) - 1;
        kt1[d] = ld_pinned(P.dim[d].knots + c);
        kt2[d] = ld_pinned(P.dim[d].knots + c + 1);
        kr[d] = ld_pinned(P.dim[d].recip + c);
    }
    const int cnt = (int)(s1 - s0);
    uint32_t ko[PT];
#pragma unroll
    for (int q = 0; q < PT; ++q) ko[q] = s0 + q;
    if (P.perm) {
#pragma unroll
        for (int q = 0; q < PT; ++q) ko[q] = ld_pinned(P.perm + s0 + (q < cnt ? q : cnt - 1));
    }
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<float>& D = P.dim[d];
        const float t1 = kt1[d], t2 = kt2[d], r = kr[d];
        const int order = DER == -2 ? D.order : (DER == d ? 1 : 0);
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            if (order == 0) { w[q][d][0] = (t2 - xs[q][d]) * r; w[q][d][1] = (xs[q][d] - t1) * r; }
            else { w[q][d][0] = -r; w[q][d][1] = r; }
        }
    }
    float2 lo[PT], hi[PT];
    LerpT<N - 1, PT>::run(ub, w, stride, lo, hi);
#pragma unroll
    for (int q = 0; q < PT; ++q) {
        if (q >= cnt) break;
        if (P.staged_records) {
            *reinterpret_cast<float4*>(P.out + (int64_t)ko[q] * P.out_point_stride) = make_float4(lo[q].x, lo[q].y, hi[q].x, hi[q].y);
        } else {
            float* o = P.out + (int64_t)ko[q] * P.out_point_stride;
            __stcs(o, lo[q].x);
            __stcs(o + P.out_comp_stride, lo[q].y);
            __stcs(o + 2 * P.out_comp_stride, hi[q].x);
            __stcs(o + 3 * P.out_comp_stride, hi[q].y);
        }
    }
}

template <int N, int PT, int MINB>
cudaError_t launch_linear_tile(const EvalParams<float>& P, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_items + TBLOCK - 1) / TBLOCK);
    const int pf = (P.knobs && P.knobs->pf1 >= 0) ? P.knobs->pf1 : 2 * 148 * MINB;
    int der = -1;
    for (int d = 0; d < N; ++d) {
        if (P.dim[d].order == 1 && der == -1) der = d;
        else if (P.dim[d].order != 0) der = -2;
    }
    if (der == -1) linear_tile_kernel<N, PT, MINB, -1><<<nb, TBLOCK, 0, s>>>(P, pf);
    else linear_tile_kernel<N, PT, MINB, -2><<<nb, TBLOCK, 0, s>>>(P, pf);
    return cudaGetLastError();
}

// item-major copies of the plan's inputs (see unstructured_tile_kernel)
template <typename T, int N>
__global__ void __launch_bounds__(256) plan_item_layout_kernel(const __grid_constant__ EvalParams<T> P, int pts, uint32_t* __restrict__ item_cells,
                                                               T* t0, T* t1, T* t2, T* t3, T* t4, T* t5) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= P.n_items) return;
    const uint32_t s0 = P.items[i], s1 = P.items[i + 1];
    const int cnt = (int)(s1 - s0);
    item_cells[i] = P.cells[s0];
    T* dst[6] = {t0, t1, t2, t3, t4, t5};
    for (int q = 0; q < pts; ++q) {
        const int64_t k = (int64_t)s0 + (q < cnt ? q : cnt - 1);
#pragma unroll
        for (int d = 0; d < N; ++d) dst[d][i * pts + q] = __ldcs(P.dim[d].t_eval + k);
    }
}

template <typename T, int N, int W, int C, int PT, int MINB>
cudaError_t launch_tile(const EvalParams<T>& P, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_items + TBLOCK - 1) / TBLOCK);
    // L2 prefetch distance in CTAs: about twice the CTAs resident on the device (148 SMs x MINB)
    const int pf = (P.knobs && P.knobs->pf1 >= 0) ? P.knobs->pf1 : 2 * 148 * MINB;
    int der = -1;
    for (int d = 0; d < N; ++d) {
        if (P.dim[d].order == 1 && der == -1) der = d;
        else if (P.dim[d].order != 0) der = -2;
    }
    if constexpr (N == 3 && W == 4 && sizeof(T) == 8 && (C == 3 || C == 4)) {
        // stencils staged in shared memory where the plan is dense enough for a warp's items to share a few cells:
        // >= 8 items per cell on average (<= 5 cells per warp, TS_SLOTS = 6)
        if (tile_smem_wanted<T, N, W>(P)) {
            const size_t smem = TileSmem<T, C>::BYTES;
#define DIND_TILE_SM(DD)                                                                                                              \
    do {                                                                                                                              \
        cudaError_t e_ = cudaFuncSetAttribute(unstructured_tile_smem_kernel<T, C, PT, MINB, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e_ != cudaSuccess) return e_;                                                                                             \
        unstructured_tile_smem_kernel<T, C, PT, MINB, DD><<<nb, TBLOCK, smem, s>>>(P, pf);                                            \
    } while (0)
            if (der == -1) DIND_TILE_SM(-1);
            else if (der == 0) DIND_TILE_SM(0);
            else if (der == 1) DIND_TILE_SM(1);
            else if (der == 2) DIND_TILE_SM(2);
            else DIND_TILE_SM(-2);
#undef DIND_TILE_SM
            return cudaGetLastError();
        }
    }
#define DIND_TILE(DD) unstructured_tile_kernel<T, N, W, C, PT, MINB, DD><<<nb, TBLOCK, 0, s>>>(P, pf)
    if (der == -1) DIND_TILE(-1);
    else if (der == 0) DIND_TILE(0);
    else if (der == 1 && N > 1) DIND_TILE((N > 1 ? 1 : -2));
    else if (der == 2 && N > 2) DIND_TILE((N > 2 ? 2 : -2));
    else DIND_TILE(-2);
#undef DIND_TILE
    return cudaGetLastError();
}

}  // namespace

// Points per item for this configuration (0 = the one-point-per-thread kernels stay in charge).  Worth it where the
// stencil is large against the per-point inputs: B-spline stencils (W >= 3) of vector-valued u.
template <typename T>
int tile_points_per_item(const EvalParams<T>& P, bool gradient) {
    if (!P.cells || !aos_supported<T>(P) || P.n_points >= (1LL << 31)) return 0;
    const int N = P.n_in, W = P.dim[0].width, C = P.n_comp;
    const int forced = P.knobs ? P.knobs->c3_tile : -1;
    if (forced == 0) return 0;
    if (gradient) return (sizeof(T) == 8 && N == 3 && W == 4 && C == 3) ? 2 : 0;   // 12 accumulators per point: two points per item
    // 3 points per item (168 registers, 3 CTAs/SM): 29.4 Gpoints/s on config 3; 2 per item 28.6, 4 per item (254
    // registers) 16.3 — DESIGN.md §4.  DIND_C3_TILE=2 keeps the 2-point variant reachable for experiments.
    if (sizeof(T) == 8 && N == 3 && W == 4 && C == 3) return forced == 2 ? 2 : 3;
    if (sizeof(T) == 4 && N == 5 && W == 2 && C == 4) {
        bool linear = true;
        double cells_total = 1.0;
        for (int d = 0; d < N; ++d) {
            linear = linear && P.dim[d].kind == LINEAR;
            cells_total *= (double)(P.dim[d].n_u - 1);
        }
        if (!linear) return 0;
        if (forced == 4) return 4;   // experiment knob DIND_C3_TILE=4
        // expected points per cell for points spread over the table: hardly any -> no items (one point per thread).
        // 2 points per item at 8 CTAs/SM beats 4 per item (166 registers, 3 CTAs/SM) at every density measured:
        // 1e8 points (3.5 per cell) 57.0 vs 43.9 Gpoints/s, 5e8 points (17 per cell) 66.6 vs 54.0 (63.6 at 4 CTAs/SM)
        const double density = (double)P.n_points / cells_total;
        return density >= 1.0 ? 2 : 0;
    }
    return 0;
}

template <typename T>
bool try_eval_fast_tile(const EvalParams<T>& P, bool gradient, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    if (!P.items || !P.u_aos) return false;
    const int N = P.n_in, W = P.dim[0].width, C = P.n_comp, PT = P.item_pts;
    *launches = 1;
    if (gradient) {
        if constexpr (sizeof(T) == 8) {
            if (N == 3 && W == 4 && C == 3 && PT == 2) {
                *err = launch_gradient_tile<T, 3, 4, 3, 2, 2>(P, s);
                *name = "unstructured_gradient_tile<N=3,W=4,C=3,PT=2>";
                return true;
            }
        }
        return false;
    }
    if constexpr (sizeof(T) == 4) {
        if (N == 5 && W == 2 && C == 4) {
            // 8 CTAs/SM (64 registers, no spills): the kernel waits on gathers, resident warps are what hides them
            // (plan order, 1e8 points: 5 CTAs 52.4, 6 CTAs 53.6, 8 CTAs 57.0 Gpoints/s)
            if (PT == 2) { *err = launch_linear_tile<5, 2, 8>(P, s); *name = "linear_tile<N=5,C=4,PT=2,ffma2>"; return true; }
            if (PT == 4) { *err = launch_linear_tile<5, 4, 4>(P, s); *name = "linear_tile<N=5,C=4,PT=4,ffma2>"; return true; }
        }
    }
    if constexpr (sizeof(T) == 8) {
        if (N == 3 && W == 4 && C == 3) {
            if (PT == 2) { *err = launch_tile<T, 3, 4, 3, 2, 4>(P, s); *name = "unstructured_tile<N=3,W=4,C=3,PT=2>"; return true; }
            if (PT == 3) { *err = launch_tile<T, 3, 4, 3, 3, 3>(P, s); *name = "unstructured_tile<N=3,W=4,C=3,PT=3>"; return true; }
        }
    }
    return false;
}

// P: plan-order parameters with P.items / P.n_items / P.cells set and dim[d].t_eval = the sorted coordinates
template <typename T>
cudaError_t plan_item_layout(const EvalParams<T>& P, int pts, uint32_t* item_cells, T* const* item_t, cudaStream_t s) {
    const unsigned nb = (unsigned)((P.n_items + 255) / 256);
    T* t[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int d = 0; d < P.n_in && d < 6; ++d) t[d] = item_t[d];
    switch (P.n_in) {
        case 1: plan_item_layout_kernel<T, 1><<<nb, 256, 0, s>>>(P, pts, item_cells, t[0], t[1], t[2], t[3], t[4], t[5]); break;
        case 2: plan_item_layout_kernel<T, 2><<<nb, 256, 0, s>>>(P, pts, item_cells, t[0], t[1], t[2], t[3], t[4], t[5]); break;
        case 3: plan_item_layout_kernel<T, 3><<<nb, 256, 0, s>>>(P, pts, item_cells, t[0], t[1], t[2], t[3], t[4], t[5]); break;
        case 4: plan_item_layout_kernel<T, 4><<<nb, 256, 0, s>>>(P, pts, item_cells, t[0], t[1], t[2], t[3], t[4], t[5]); break;
        case 5: plan_item_layout_kernel<T, 5><<<nb, 256, 0, s>>>(P, pts, item_cells, t[0], t[1], t[2], t[3], t[4], t[5]); break;
        case 6: plan_item_layout_kernel<T, 6><<<nb, 256, 0, s>>>(P, pts, item_cells, t[0], t[1], t[2], t[3], t[4], t[5]); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
template cudaError_t plan_item_layout<float>(const EvalParams<float>&, int, uint32_t*, float* const*, cudaStream_t);
template cudaError_t plan_item_layout<double>(const EvalParams<double>&, int, uint32_t*, double* const*, cudaStream_t);

#define DIND_INST(T)                                                      \
    template int tile_points_per_item<T>(const EvalParams<T>&, bool);     \
    template bool try_eval_fast_tile<T>(const EvalParams<T>&, bool, cudaStream_t, const char**, cudaError_t*, int*);
DIND_INST(float)
DIND_INST(double)

}  // namespace dind
