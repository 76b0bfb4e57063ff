// dind_fast_unstructured_rec.cu — eval_unstructured! for multilinear, vector-valued lookup tables that do not
// fit in L2, at iid points (BASELINE config 5: u = 32^5 x 4 Float32 = 537 MB, 4e9 random points).
//
// What bounds a random gather on B200 (scripts/microbench_gather.cu, profiles/r2b_microbench_gather.txt): an SM
// keeps a fixed number of L1 misses in flight, whatever their size.  A warp instruction whose 32 lanes ask for 32
// different 32-byte sectors moves 1.5 TB/s from a table in HBM; the same bytes requested as whole 128-byte lines
// (4 lanes x 32 bytes) move 7.0 TB/s.  With u in the reference's layout (or its component-fastest copy) the 2^N
// corners of a cell are 2^(N-1) row segments of two nodes in 2^(N-1) distant places — no request can be a whole
// line, and ncu shows 1.8 KB of DRAM reads per point for 512 B of corner data.  B200 has the memory to trade for
// it: the handle keeps a CELL-RECORD copy of u,
//
//     u_rec[cell(i_2..i_N)][i_1][slot(k_2..k_N)][component]        cell over the n_d - 1 intervals of axes 2..N
//
// in which the 2^(N-1) corner nodes of a cell at one i_1 are contiguous, and so — i_1 and i_1 + 1 being
// neighbours — all 2^N corners of a point are ONE aligned, contiguous block (config 5: 512 B = four 128-byte
// lines).  Size: prod(n_d - 1, d >= 2) * n_1 * 2^(N-1) nodes (config 5: 7.6 GB of the 180 GB), built once per
// dind_set_u by one pass over the component-fastest copy.
//
// FOUR LANES gather and contract one point (the search and the weights were computed one point per lane and
// arrive by shuffle): every load instruction of a quad covers one whole line of the block (lane q its q-th 32 bytes).  The slots of a half-record are ordered so that what lane q receives is the sub-cube of axes
// 1..N-2 at (k_{N-1}, k_N) = q: the lane contracts those axes in registers, the last two axes are two
// butterfly exchanges (shfl.xor 1, 2) after which every lane of the quad holds the result and lane c stores
// component c.  The weights and the order of the FMAs are those of unstructured_aos_kernel (each partial sum is
// formed by the same fma chain, only on another lane), so the results are bit-identical to it
// (tests/test_gpu_parity.py::test_cell_record_table_is_bit_identical).
//
// Replaces: eval_kernel + _interpolate! for LinearInterpolationDimension (src/interpolation_parallel.jl:105-138,
// src/interpolation_methods.jl:1-36); the reference has one layout for u and no counterpart of this table.
#include <type_traits>

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int RBLOCK = 256;

// the Linear stencil of one dimension — same arithmetic as stencil_ct in dind_fast_unstructured_aos.cu
template <typename T>
__device__ __forceinline__ int linear_stencil(const DimParams<T>& D, T x, int cached_idx, T (&w)[2]) {
    const int n = D.n_knots;
    const int idx = cached_idx ? cached_idx : clampi(dim_count<T, false>(D, x, to_key(x)), 1, n - 1);
    const int c = idx - 1;
    const T t1 = __ldg(D.knots + c), t2 = __ldg(D.knots + c + 1), r = __ldg(D.recip + c);
    if (D.order == 0) { w[0] = (t2 - x) * r; w[1] = (x - t1) * r; }
    else { w[0] = -r; w[1] = r; }
    return c;
}

// 32 bytes that are used once: no L1 allocation (the knot / search tables stay resident there)
__device__ __forceinline__ void ld32(const void* p, float (&q)[8]) {
    asm("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(q[0]), "=f"(q[1]), "=f"(q[2]), "=f"(q[3]), "=f"(q[4]), "=f"(q[5]), "=f"(q[6]), "=f"(q[7]) : "l"(p));
}
__device__ __forceinline__ void ld32(const void* p, double (&q)[4]) {
    asm("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(q[0]), "=d"(q[1]), "=d"(q[2]), "=d"(q[3]) : "l"(p));
}

// slot of corner (k_2..k_N) inside a half-record, in nodes.  A lane's 32-byte chunk holds NPC = 32 / node bytes
// nodes; chunk 4 * jj + q belongs to lane q; (jj, r) enumerate the corners of axes 2..N-2 (axis 2 fastest) and
// q = k_{N-1} + 2 k_N.  `corner` has bit d-2 for axis d (1-based).
__host__ __device__ __forceinline__ int rec_slot_to_corner(int slot, int npc, int n_in) {
    const int chunk = slot / npc, r = slot % npc;
    return ((chunk >> 2) * npc + r) | ((chunk & 3) << (n_in - 3));
}

template <typename T, int C> struct RecGeom {
    static constexpr int S = C == 3 ? 4 : C;
    static constexpr int NODE_BYTES = S * (int)sizeof(T);
    static constexpr int NPC = 32 / NODE_BYTES;        // nodes per 32-byte chunk
};

// one 32-byte chunk -> NPC nodes
template <typename T, int C>
__device__ __forceinline__ void load_chunk(const char* __restrict__ p, Vec<T, C>* nd) {
    constexpr int S = RecGeom<T, C>::S, NPC = RecGeom<T, C>::NPC;
    if constexpr (sizeof(T) == 4) {
        float q[8];
        ld32(p, q);
#pragma unroll
        for (int r = 0; r < NPC; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) nd[r].v[c] = (T)q[r * S + c];
    } else {
        double q[4];
        ld32(p, q);
#pragma unroll
        for (int r = 0; r < NPC; ++r)
#pragma unroll
            for (int c = 0; c < C; ++c) nd[r].v[c] = (T)q[r * S + c];
    }
}

// the lane's sub-cube: axes 1..D+1 (0-based 0..D) at local corner offset OFF — the fma chains of ContractV
template <typename T, int D, int N, int C, int M, int OFF>
struct QuadContract {
    static __device__ __forceinline__ Vec<T, C> run(const Vec<T, C> (&nd)[2][M], const T (&w)[N][2]) {
        const Vec<T, C> t0 = QuadContract<T, D - 1, N, C, M, OFF>::run(nd, w);
        const Vec<T, C> t1 = QuadContract<T, D - 1, N, C, M, OFF + (1 << (D - 1))>::run(nd, w);
        Vec<T, C> acc;
#pragma unroll
        for (int c = 0; c < C; ++c) acc.v[c] = fma(w[D][1], t1.v[c], fma(w[D][0], t0.v[c], T(0)));
        return acc;
    }
};
template <typename T, int N, int C, int M, int OFF>
struct QuadContract<T, 0, N, C, M, OFF> {
    static __device__ __forceinline__ Vec<T, C> run(const Vec<T, C> (&nd)[2][M], const T (&w)[N][2]) {
        Vec<T, C> acc;
#pragma unroll
        for (int c = 0; c < C; ++c) acc.v[c] = fma(w[0][1], nd[1][OFF].v[c], fma(w[0][0], nd[0][OFF].v[c], T(0)));
        return acc;
    }
};

// axis `d` lives on bit `bit` of the quad lane: both lanes form fma(w[1], t(k=1), fma(w[0], t(k=0), 0))
template <typename T, int C>
__device__ __forceinline__ Vec<T, C> quad_axis(const Vec<T, C>& mine, const T (&w)[2], int q, int bit) {
    Vec<T, C> acc;
    const bool hi = (q >> bit) & 1;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const T other = __shfl_xor_sync(0xffffffffu, mine.v[c], 1 << bit);
        const T t0 = hi ? other : mine.v[c], t1 = hi ? mine.v[c] : other;
        acc.v[c] = fma(w[1], t1, fma(w[0], t0, T(0)));
    }
    return acc;
}

// A warp owns 32 consecutive points.  Phase A, one point per lane: interval search and weights of every dimension
// (t_eval loads coalesced, no redundant work).  Phase B, four rounds: quad g gathers and contracts point 8r + g, whose
// weights and record index arrive by shuffle from the lane that computed them.
// 4 CTAs/SM = 64 registers: ptxas then keeps a second round's loads in flight (config 5, 1e8 iid points: 12.2
// Gpoints/s = 6.7 TB/s of record traffic; 52 registers without the bound 10.2, 5 / 6 CTAs per SM 10.7 / 10.6)
template <typename T, int N, int C>
__global__ void __launch_bounds__(RBLOCK, 4) unstructured_rec_kernel(const __grid_constant__ EvalParams<T> P) {
    constexpr int NPC = RecGeom<T, C>::NPC;
    constexpr int M = 1 << (N - 3);            // the lane's nodes per i_1
    constexpr int JJ = M / NPC;                // its chunks per half-record
    constexpr int HALF_BYTES = (1 << (N - 1)) * RecGeom<T, C>::NODE_BYTES;
    static_assert(JJ >= 1, "a lane's share of a line must be whole nodes of one half-record");
    const int64_t k0 = (int64_t)blockIdx.x * RBLOCK + (threadIdx.x & ~31);   // the warp's first point
    const int lane = threadIdx.x & 31, q = lane & 3;
    const int64_t k = min(k0 + lane, P.n_points - 1);
    T w[N][2];
    int64_t rec = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const DimParams<T>& D = P.dim[d];
        const T x = __ldcs(D.t_eval + k);
        const int cached = P.use_cached_idx ? __ldcs(D.idx_eval + k) : 0;
        rec += (int64_t)linear_stencil<T>(D, x, cached, w[d]) * D.rec_stride;
    }
    const char* mine = reinterpret_cast<const char*>(P.u_rec) + rec * HALF_BYTES;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int src = 8 * r + (lane >> 2);
        const char* base = reinterpret_cast<const char*>(__shfl_sync(0xffffffffu, (unsigned long long)mine, src)) + q * 32;
        Vec<T, C> nd[2][M];
#pragma unroll
        for (int k1 = 0; k1 < 2; ++k1)
#pragma unroll
            for (int jj = 0; jj < JJ; ++jj) load_chunk<T, C>(base + k1 * HALF_BYTES + jj * 128, &nd[k1][jj * NPC]);
        T wr[N][2];
#pragma unroll
        for (int d = 0; d < N; ++d) {
            wr[d][0] = __shfl_sync(0xffffffffu, w[d][0], src);
            wr[d][1] = __shfl_sync(0xffffffffu, w[d][1], src);
        }
        Vec<T, C> acc = QuadContract<T, N - 3, N, C, M, 0>::run(nd, wr);
        acc = quad_axis<T, C>(acc, wr[N - 2], q, 0);
        acc = quad_axis<T, C>(acc, wr[N - 1], q, 1);
        // lane c stores component c: a warp writes 8 consecutive points of every component column
        const int64_t kp = k0 + src;
        if (kp < P.n_points) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (q == c) __stcs(P.out + (int64_t)c * P.out_comp_stride + kp * P.out_point_stride, acc.v[c]);
        }
    }
}

struct RecShape {
    int n_in, S;
    int n1;               // nodes along axis 1
    int cells[MAXN];      // n_d - 1 for d >= 1 (0-based), [0] unused
    int64_t stride[MAXN]; // node stride of dimension d in u
    int64_t n_nodes;      // nodes of the record table
};

// one thread per node of the record table (writes coalesced; the reads of a CTA fall into 2^(N-1) row segments)
// slot -> corner: rec_slot_to_corner
template <typename T, int S>
__global__ void __launch_bounds__(RBLOCK) rec_build_kernel(const T* __restrict__ u_aos, T* __restrict__ u_rec, const RecShape sh) {
    const int64_t e = (int64_t)blockIdx.x * RBLOCK + threadIdx.x;
    if (e >= sh.n_nodes) return;
    const int nc = 1 << (sh.n_in - 1);
    const int corner = rec_slot_to_corner((int)(e & (nc - 1)), 32 / (S * (int)sizeof(T)), sh.n_in);
    int64_t r = e >> (sh.n_in - 1);
    int64_t src = r % sh.n1;
    r /= sh.n1;
    for (int d = 1; d < sh.n_in; ++d) {
        const int i = (int)(r % sh.cells[d]);
        r /= sh.cells[d];
        src += (int64_t)(i + ((corner >> (d - 1)) & 1)) * sh.stride[d];
    }
    using V = typename std::conditional<S * sizeof(T) == 8, float2, float4>::type;
    constexpr int NV = (int)(S * sizeof(T) / sizeof(V));
    const V* s = reinterpret_cast<const V*>(u_aos + src * S);
    V* o = reinterpret_cast<V*>(u_rec + e * S);
#pragma unroll
    for (int j = 0; j < NV; ++j) __stcs(o + j, __ldg(s + j));
}

template <typename T>
RecShape rec_shape(const EvalParams<T>& P) {
    RecShape sh{};
    sh.n_in = P.n_in;
    sh.S = aos_node_stride(P.n_comp);
    sh.n1 = P.dim[0].n_u;
    sh.n_nodes = (int64_t)sh.n1 << (P.n_in - 1);
    for (int d = 1; d < P.n_in; ++d) {
        sh.cells[d] = P.dim[d].n_u - 1;
        sh.stride[d] = P.dim[d].u_stride;
        sh.n_nodes *= sh.cells[d];
    }
    return sh;
}

}  // namespace

template <typename T>
bool rec_supported(const EvalParams<T>& P) {
    if (!aos_supported<T>(P) || P.n_in < 3 || P.n_in > 6 || P.coherent_points) return false;
    // a lane's share of a line (32 bytes) must be whole nodes of one half-record: 2^(N-1) nodes >= 128 bytes
    if ((size_t)(1 << (P.n_in - 1)) * aos_node_stride(P.n_comp) * sizeof(T) < 128) return false;
    double nodes = (double)P.dim[0].n_u * (double)(1 << (P.n_in - 1));
    for (int d = 0; d < P.n_in; ++d) {
        if (P.dim[d].kind != LINEAR || P.dim[d].width != 2 || P.dim[d].n_u < 2) return false;
        if (d) nodes *= (double)(P.dim[d].n_u - 1);
    }
    return nodes < 4.0e18 / 64;
}

template <typename T>
size_t rec_table_bytes(const EvalParams<T>& P) {
    return (size_t)rec_shape<T>(P).n_nodes * aos_node_stride(P.n_comp) * sizeof(T);
}

// record stride of every dimension (in records of 2^(N-1) nodes): axis 1 is 1, axis d >= 2 walks the cells
template <typename T>
void rec_fill_strides(EvalParams<T>& P) {
    int64_t s = P.dim[0].n_u;
    P.dim[0].rec_stride = 1;
    for (int d = 1; d < P.n_in; ++d) {
        P.dim[d].rec_stride = s;
        s *= P.dim[d].n_u - 1;
    }
}

template <typename T>
cudaError_t launch_rec_build(const EvalParams<T>& P, T* u_rec, cudaStream_t s) {
    const RecShape sh = rec_shape<T>(P);
    const int64_t nb = (sh.n_nodes + RBLOCK - 1) / RBLOCK;
    if (nb >= (1LL << 31)) return cudaErrorInvalidValue;
    if (sh.S == 2) rec_build_kernel<T, 2><<<(unsigned)nb, RBLOCK, 0, s>>>(P.u_aos, u_rec, sh);
    else rec_build_kernel<T, 4><<<(unsigned)nb, RBLOCK, 0, s>>>(P.u_aos, u_rec, sh);
    return cudaGetLastError();
}

template <typename T>
bool try_eval_fast_unstructured_rec(const EvalParams<T>& P, cudaStream_t s, const char** name, cudaError_t* err, int* launches) {
    if (!P.u_rec || !rec_supported<T>(P)) return false;
    const unsigned nb = (unsigned)((P.n_points + RBLOCK - 1) / RBLOCK);
    const int N = P.n_in, C = P.n_comp;
    constexpr int TB = (int)sizeof(T);
    *launches = 1;
#define DIND_REC(NN, CC)                                                   \
    if constexpr ((1 << (NN - 1)) * (CC == 3 ? 4 : CC) * TB >= 128)        \
    if (N == NN && C == CC) {                                              \
        unstructured_rec_kernel<T, NN, CC><<<nb, RBLOCK, 0, s>>>(P);       \
        *err = cudaGetLastError();                                         \
        *name = "unstructured_rec<N=" #NN ",C=" #CC ">";                   \
        return true;                                                       \
    }
#define DIND_REC_N(NN) DIND_REC(NN, 2) DIND_REC(NN, 3) DIND_REC(NN, 4)
    DIND_REC_N(3) DIND_REC_N(4) DIND_REC_N(5) DIND_REC_N(6)
#undef DIND_REC_N
#undef DIND_REC
    return false;
}

#define DIND_INST(T)                                                                                              \
    template bool rec_supported<T>(const EvalParams<T>&);                                                         \
    template size_t rec_table_bytes<T>(const EvalParams<T>&);                                                     \
    template void rec_fill_strides<T>(EvalParams<T>&);                                                            \
    template cudaError_t launch_rec_build<T>(const EvalParams<T>&, T*, cudaStream_t);                             \
    template bool try_eval_fast_unstructured_rec<T>(const EvalParams<T>&, cudaStream_t, const char**, cudaError_t*, int*);
DIND_INST(float)
DIND_INST(double)

}  // namespace dind
