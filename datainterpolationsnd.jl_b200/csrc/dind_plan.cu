// dind_plan.cu — cell-sorted evaluation plan for eval_unstructured.
//
// The reference caches, per t_eval, the interval index and the basis values of every point
// (src/interpolation_utils.jl:143-161, src/spline_utils.jl:164-191) so that re-evaluation with a
// new `u` or another derivative order is cheap.  The B200 analogue of that cache is an ORDER:
// with iid points every lane of a warp gathers its own (degree+1)^N stencil (config 3: 1.5 KB per
// point through L1/L2; config 5: ~6 KB per point of HBM sector traffic), whereas points of the
// same cell read the same addresses.  The plan sorts the point indices once by the offset of
// their first stencil node in `u` (radix sort, CUB — cache construction, not the hot path) and
// keeps sorted copies of t_eval; the evaluation kernels then run over the sorted list with
// warp-coherent gathers and scatter each result to its original position.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int BLOCK = 256;

// PACKED: the key is the per-dimension first nodes in bit fields (last dimension most significant: the same
// order as the offset in u) — kept after the sort as the plan's cell words, which the evaluation kernels decode
// instead of searching again.
// recs != nullptr: the coordinates of point k are also written as one record of rec_elems elements (a multiple of
// 32 bytes), so that the sorted copies of t_eval can be produced by ONE random read per point (plan_gather_rec_kernel)
// instead of one per point and dimension: a random 4- or 8-byte read costs a whole L1 miss slot (config 5, 1e8 points:
// five gathers of 1.77 ms each).
template <typename T, bool PACKED>
__global__ void __launch_bounds__(BLOCK) plan_key_kernel(const __grid_constant__ EvalParams<T> P, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals, T* __restrict__ recs, int rec_elems) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    int64_t base = 0;
    for (int d = 0; d < P.n_in; ++d) {
        const DimParams<T>& D = P.dim[d];
        const T x = __ldcs(D.t_eval + k);
        if (recs) recs[k * rec_elems + d] = x;
        const int i0 = dim_first_node<T>(D, x);
        if (PACKED) base |= (int64_t)i0 << D.cell_shift;
        else base += (int64_t)i0 * D.u_stride;
    }
    keys[k] = (uint32_t)base;
    vals[k] = (uint32_t)k;
}

struct GatherDst {
    void* p[MAXN];
};
// t_sorted[d][k] = recs[perm[k]][d]: the record is read with 32-byte loads, the N_in output streams are coalesced
template <typename T, int N>
__global__ void __launch_bounds__(BLOCK) plan_gather_rec_kernel(const T* __restrict__ recs, const uint32_t* __restrict__ perm,
                                                                const GatherDst dst, int64_t n) {
    constexpr int PER = 32 / (int)sizeof(T);            // elements per 32-byte load
    constexpr int NL = (N + PER - 1) / PER;
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    const T* r = recs + (int64_t)__ldcs(perm + k) * (NL * PER);
    T v[NL * PER];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        if constexpr (sizeof(T) == 4) {
            asm("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                : "=f"(v[j * 8]), "=f"(v[j * 8 + 1]), "=f"(v[j * 8 + 2]), "=f"(v[j * 8 + 3]), "=f"(v[j * 8 + 4]), "=f"(v[j * 8 + 5]),
                  "=f"(v[j * 8 + 6]), "=f"(v[j * 8 + 7]) : "l"(r + j * 8));
        } else {
            asm("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];"
                : "=d"(v[j * 4]), "=d"(v[j * 4 + 1]), "=d"(v[j * 4 + 2]), "=d"(v[j * 4 + 3]) : "l"(r + j * 4));
        }
    }
#pragma unroll
    for (int d = 0; d < N; ++d) __stcs((T*)dst.p[d] + k, v[d]);
}

template <typename T>
__global__ void __launch_bounds__(BLOCK) plan_gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ perm,
                                                            T* __restrict__ dst, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k < n) dst[k] = src[perm[k]];
}

// ---- un-permutation through window-ordered staging slots (records smaller than a sector) ------------------------------------------------------------------
// Bringing plan-ordered results back to the caller's order is a random permutation over the whole
// output (config 3: 480 MB, config 5: 1.6 GB): as one gather it runs at DRAM-sector/TLB speed.
// Instead the original index space is cut into windows of 2^shift points; the evaluation kernel
// writes the result of plan position k to staging slot slot2[k], where the slots of a window are handed
// out in arrival order (bucket_cursor_kernel below) — every window's region fills front to back while the
// kernel sweeps the plan, sequential write streams that L2 write-combines — and a second pass copies slot j
// to its destination dest2[j], which lies inside a window small enough to be assembled in shared memory
// (plan_unbucket_kernel).  Used for records smaller than a 32-byte sector; whole-sector records are staged
// at their original index (plan_unstage below).
// Second pass: one CTA per window of UNB_W original points.  The window's records are exactly the staging
// slots [w * UNB_W, w * UNB_W + size) (the map is a permutation, so every window receives as many records
// as it has points); they are read in order, dropped at their place in a shared-memory image of the
// window (component-major), and the image leaves with fully coalesced streaming stores.
constexpr int UNB_W = 1 << PLAN_WINDOW_SHIFT;
constexpr int UNB_BLOCK = 512;

template <typename T>
__global__ void __launch_bounds__(UNB_BLOCK) plan_unbucket_kernel(const T* __restrict__ staged, const uint32_t* __restrict__ dest2,
                                                                  T* __restrict__ out, int64_t n, int ct, int rec_stride, int cg) {
    extern __shared__ __align__(16) unsigned char unb_smem[];
    T* sm = reinterpret_cast<T*>(unb_smem);   // [cg][UNB_W]
    const int64_t w0 = (int64_t)blockIdx.x * UNB_W;
    const int wn = (int)(n - w0 < UNB_W ? n - w0 : UNB_W);
    for (int c0 = 0; c0 < ct; c0 += cg) {
        const int cn = ct - c0 < cg ? ct - c0 : cg;
        for (int r = threadIdx.x; r < wn; r += UNB_BLOCK) {
            const int li = (int)(__ldcs(dest2 + w0 + r) - (uint32_t)w0);
            const T* src = staged + (w0 + r) * rec_stride + c0;
            if (sizeof(T) == 8 && rec_stride == 4 && cn == ct) {          // 32-byte record: two 16-byte loads
                const double2 a = __ldcs(reinterpret_cast<const double2*>(src));
                const double2 b = __ldcs(reinterpret_cast<const double2*>(src) + 1);
                sm[li] = (T)a.x;
                if (ct > 1) sm[UNB_W + li] = (T)a.y;
                if (ct > 2) sm[2 * UNB_W + li] = (T)b.x;
                if (ct > 3) sm[3 * UNB_W + li] = (T)b.y;
            } else if (sizeof(T) == 4 && rec_stride == 4 && cn == ct) {   // 16-byte record
                const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
                sm[li] = (T)a.x;
                if (ct > 1) sm[UNB_W + li] = (T)a.y;
                if (ct > 2) sm[2 * UNB_W + li] = (T)a.z;
                if (ct > 3) sm[3 * UNB_W + li] = (T)a.w;
            } else {
                for (int cc = 0; cc < cn; ++cc) sm[cc * UNB_W + li] = __ldcs(src + cc);
            }
        }
        __syncthreads();
        for (int cc = 0; cc < cn; ++cc)
            for (int li = threadIdx.x; li < wn; li += UNB_BLOCK) __stcs(out + w0 + li + (int64_t)(c0 + cc) * n, sm[cc * UNB_W + li]);
        __syncthreads();
    }
}

}  // namespace

// ---- work items: <= pts consecutive points of one cell -----------------------------------------------
namespace {
struct RunHead {   // k -> k if a run of equal cell words starts at k, else 0 (max-scan gives every k its run start)
    const uint32_t* cells;
    __device__ __forceinline__ uint32_t operator()(uint32_t k) const { return (k == 0 || cells[k] != cells[k - 1]) ? k : 0u; }
};
struct IsItem {
    const uint32_t* run_start;
    uint32_t pts;
    __device__ __forceinline__ bool operator()(uint32_t k) const { return (k - run_start[k]) % pts == 0; }
};
struct IsItemCount {
    IsItem f;
    __device__ __forceinline__ unsigned long long operator()(uint32_t k) const { return f(k) ? 1ull : 0ull; }
};
__global__ void items_tail_kernel(uint32_t* items, const unsigned long long* count, uint32_t n) { items[*count] = n; }
}  // namespace

cudaError_t plan_build_items(const uint32_t* cells, int64_t n, int pts, uint32_t* items, int64_t* n_items_out, void* scratch,
                             size_t* scratch_bytes, cudaStream_t s) {
    using Count = cub::CountingInputIterator<uint32_t>;
    const size_t arr = ((size_t)n * sizeof(uint32_t) + 255) & ~(size_t)255;
    Count iota(0);
    cub::TransformInputIterator<uint32_t, RunHead, Count> heads(iota, RunHead{cells});
    uint32_t* run_start = (uint32_t*)scratch;
    IsItem pred{run_start, (uint32_t)pts};
    cub::TransformInputIterator<unsigned long long, IsItemCount, Count> ones(iota, IsItemCount{pred});
    size_t b_scan = 0, b_red = 0, b_sel = 0;
    cudaError_t e = cub::DeviceScan::InclusiveScan(nullptr, b_scan, heads, run_start, cub::Max(), (int)n, s);
    if (e != cudaSuccess) return e;
    e = cub::DeviceReduce::Sum(nullptr, b_red, ones, (unsigned long long*)nullptr, (int)n, s);
    if (e != cudaSuccess) return e;
    e = cub::DeviceSelect::If(nullptr, b_sel, iota, (uint32_t*)nullptr, (unsigned long long*)nullptr, (int)n, pred, s);
    if (e != cudaSuccess) return e;
    const size_t cub_bytes = std::max(b_scan, std::max(b_red, b_sel));
    const size_t need = arr + 256 + cub_bytes;
    if (!scratch) {
        *scratch_bytes = need;
        return cudaSuccess;
    }
    if (*scratch_bytes < need) return cudaErrorInvalidValue;
    unsigned long long* d_count = (unsigned long long*)((char*)scratch + arr);
    void* cub_tmp = (char*)scratch + arr + 256;
    size_t tb = cub_bytes;
    if (!items) {   // pass 1: run starts + number of items
        e = cub::DeviceScan::InclusiveScan(cub_tmp, tb, heads, run_start, cub::Max(), (int)n, s);
        if (e != cudaSuccess) return e;
        tb = cub_bytes;
        e = cub::DeviceReduce::Sum(cub_tmp, tb, ones, d_count, (int)n, s);
        if (e != cudaSuccess) return e;
        unsigned long long c = 0;
        e = cudaMemcpyAsync(&c, d_count, sizeof c, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(s);
        *n_items_out = (int64_t)c;
        return e;
    }
    // pass 2 (run_start is still in scratch): the item starts, then the sentinel items[n_items] = n
    e = cub::DeviceSelect::If(cub_tmp, tb, iota, items, d_count, (int)n, pred, s);
    if (e != cudaSuccess) return e;
    items_tail_kernel<<<1, 1, 0, s>>>(items, d_count, (uint32_t)n);
    return cudaGetLastError();
}

// ---- un-permutation through a staging buffer in ORIGINAL order ---------------------------------------------------
// The evaluation kernel writes the result of plan position k as one aligned record to staged[perm[k]] (one random,
// whole-sector store per point — the same write pattern as the window-ordered slots above, without the slot schedule
// and its two index arrays); this pass reads the records in order and writes the component columns, both coalesced.
constexpr int UNS_BLOCK = 256;
template <typename T, int BYTES>
__global__ void __launch_bounds__(UNS_BLOCK) plan_unstage_vec_kernel(const T* __restrict__ staged, T* __restrict__ out, int64_t n, int ct) {
    const int64_t j = (int64_t)blockIdx.x * UNS_BLOCK + threadIdx.x;
    if (j >= n) return;
    constexpr int E = BYTES / (int)sizeof(T);
    T v[E];
    if constexpr (BYTES == 32) {
        const double2 a = __ldcs(reinterpret_cast<const double2*>(staged + j * E));
        const double2 b = __ldcs(reinterpret_cast<const double2*>(staged + j * E) + 1);
        memcpy(v, &a, 16);
        memcpy(reinterpret_cast<char*>(v) + 16, &b, 16);
    } else if constexpr (BYTES == 16) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(staged + j * E));
        memcpy(v, &a, 16);
    } else {
        const float2 a = __ldcs(reinterpret_cast<const float2*>(staged + j * E));
        memcpy(v, &a, 8);
    }
#pragma unroll
    for (int c = 0; c < E; ++c)
        if (c < ct) __stcs(out + j + (int64_t)c * n, v[c]);
}
// records of any size: a CTA's records are contiguous — coalesced copy to shared memory (row stride + 1: no bank
// conflicts in the column reads), then one coalesced store stream per component
template <typename T>
__global__ void __launch_bounds__(UNS_BLOCK) plan_unstage_smem_kernel(const T* __restrict__ staged, T* __restrict__ out, int64_t n, int ct,
                                                                     int stride) {
    extern __shared__ __align__(16) unsigned char uns_smem[];
    T* sm = reinterpret_cast<T*>(uns_smem);
    const int64_t j0 = (int64_t)blockIdx.x * UNS_BLOCK;
    const int np = (int)(n - j0 < UNS_BLOCK ? n - j0 : UNS_BLOCK);
    const T* src = staged + j0 * stride;
    for (int g = threadIdx.x; g < np * stride; g += UNS_BLOCK) {
        const int r = g / stride;
        sm[r * (stride + 1) + (g - r * stride)] = __ldcs(src + g);
    }
    __syncthreads();
    if ((int)threadIdx.x < np)
        for (int c = 0; c < ct; ++c) __stcs(out + j0 + threadIdx.x + (int64_t)c * n, sm[threadIdx.x * (stride + 1) + c]);
}
template <typename T>
__global__ void __launch_bounds__(UNS_BLOCK) plan_unstage_loop_kernel(const T* __restrict__ staged, T* __restrict__ out, int64_t n, int ct,
                                                                     int stride) {
    const int64_t j = (int64_t)blockIdx.x * UNS_BLOCK + threadIdx.x;
    if (j >= n) return;
    for (int c = 0; c < ct; ++c) __stcs(out + j + (int64_t)c * n, __ldg(staged + j * stride + c));
}

template <typename T>
cudaError_t plan_unstage(const T* staged, T* out, int64_t n, int ct, int rec_stride, cudaStream_t s) {
    const unsigned nb = (unsigned)((n + UNS_BLOCK - 1) / UNS_BLOCK);
    const int bytes = rec_stride * (int)sizeof(T);
    const size_t smem = (size_t)UNS_BLOCK * (rec_stride + 1) * sizeof(T);
    if (bytes == 32) plan_unstage_vec_kernel<T, 32><<<nb, UNS_BLOCK, 0, s>>>(staged, out, n, ct);
    else if (bytes == 16) plan_unstage_vec_kernel<T, 16><<<nb, UNS_BLOCK, 0, s>>>(staged, out, n, ct);
    else if (bytes == 8) plan_unstage_vec_kernel<T, 8><<<nb, UNS_BLOCK, 0, s>>>(staged, out, n, ct);
    else if (smem <= 48 * 1024) plan_unstage_smem_kernel<T><<<nb, UNS_BLOCK, smem, s>>>(staged, out, n, ct, rec_stride);
    else plan_unstage_loop_kernel<T><<<nb, UNS_BLOCK, 0, s>>>(staged, out, n, ct, rec_stride);
    return cudaGetLastError();
}
template cudaError_t plan_unstage<float>(const float*, float*, int64_t, int, int, cudaStream_t);
template cudaError_t plan_unstage<double>(const double*, double*, int64_t, int, int, cudaStream_t);

// The slot schedule without a sort.  perm is a permutation, so window w receives exactly its own 2^shift points and
// its slots are [w << shift, ...): plan position k takes the next free slot of its window from a cursor (one L2
// atomic on n >> shift counters).  The kernel sweeps the plan front to back, so every window still fills (nearly) in
// plan order — which is all the sequential-fill argument needs; which slot a record gets inside its window does not
// change any result (plan_unbucket places it by dest2).
__global__ void __launch_bounds__(BLOCK) bucket_cursor_kernel(const uint32_t* __restrict__ perm, uint32_t* __restrict__ cursor,
                                                              uint32_t* __restrict__ slot2, uint32_t* __restrict__ dest2, int64_t n, int shift) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= n) return;
    const uint32_t dest = __ldcs(perm + k);
    const uint32_t w = dest >> shift;
    const uint32_t j = (w << shift) + atomicAdd(cursor + w, 1u);
    __stcs(slot2 + k, j);
    dest2[j] = dest;
}

cudaError_t plan_bucket_schedule(const uint32_t* perm, uint32_t* slot2, uint32_t* dest2, int64_t n, void* scratch,
                                        size_t* scratch_bytes, cudaStream_t s) {
    const size_t need = (size_t)((n >> PLAN_WINDOW_SHIFT) + 1) * sizeof(uint32_t);
    if (!scratch) {
        *scratch_bytes = need;
        return cudaSuccess;
    }
    if (*scratch_bytes < need) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(scratch, 0, need, s);
    if (e != cudaSuccess) return e;
    bucket_cursor_kernel<<<(unsigned)((n + BLOCK - 1) / BLOCK), BLOCK, 0, s>>>(perm, (uint32_t*)scratch, slot2, dest2, n, PLAN_WINDOW_SHIFT);
    return cudaGetLastError();
}

template <typename T>
cudaError_t plan_unbucket(const T* staged, const uint32_t* dest2, T* out, int64_t n, int ct, int rec_stride, cudaStream_t s) {
    const int cg_max = (96 * 1024) / (UNB_W * (int)sizeof(T));   // components per round: <= 96 KB of shared memory
    const int cg = ct < cg_max ? ct : cg_max;
    const size_t smem = (size_t)cg * UNB_W * sizeof(T);
    // per device (a process may hold handles on several devices), and cheap: set it on every launch
    cudaError_t e = cudaFuncSetAttribute(plan_unbucket_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    plan_unbucket_kernel<T><<<(unsigned)((n + UNB_W - 1) / UNB_W), UNB_BLOCK, smem, s>>>(staged, dest2, out, n, ct, rec_stride, cg);
    return cudaGetLastError();
}
template cudaError_t plan_unbucket<float>(const float*, const uint32_t*, float*, int64_t, int, int, cudaStream_t);
template cudaError_t plan_unbucket<double>(const double*, const uint32_t*, double*, int64_t, int, int, cudaStream_t);

template <typename T>
int plan_rec_elems(int n_in) {
    const int per = 32 / (int)sizeof(T);
    return (n_in + per - 1) / per * per;
}
template int plan_rec_elems<float>(int);
template int plan_rec_elems<double>(int);

template <typename T>
cudaError_t plan_sort_points(const EvalParams<T>& P, uint32_t* perm, T* const* t_sorted, uint32_t* cells, int cell_bits,
                             void* scratch, size_t* scratch_bytes, T* recs, cudaStream_t s) {
    const int64_t n = P.n_points;
    int end_bit = 1;
    while (end_bit < 32 && (1LL << end_bit) < P.u_comp_stride) ++end_bit;
    if (cells) end_bit = cell_bits;
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, n, 0, end_bit, s);
    if (e != cudaSuccess) return e;
    const size_t arr = ((size_t)n * sizeof(uint32_t) + 255) & ~(size_t)255;
    const size_t need = 3 * arr + cub_bytes;   // keys_in, keys_out, vals_in (+ CUB temp); vals_out = perm
    if (!scratch) {
        *scratch_bytes = need;
        return cudaSuccess;
    }
    if (*scratch_bytes < need) return cudaErrorInvalidValue;
    char* base = (char*)scratch;
    uint32_t* keys_in = (uint32_t*)base;
    uint32_t* keys_out = (uint32_t*)(base + arr);
    uint32_t* vals_in = (uint32_t*)(base + 2 * arr);
    void* cub_tmp = base + 3 * arr;
    const unsigned nb = (unsigned)((n + BLOCK - 1) / BLOCK);
    const int rec_elems = plan_rec_elems<T>(P.n_in);
    if (cells) plan_key_kernel<T, true><<<nb, BLOCK, 0, s>>>(P, keys_in, vals_in, recs, rec_elems);
    else plan_key_kernel<T, false><<<nb, BLOCK, 0, s>>>(P, keys_in, vals_in, recs, rec_elems);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, cells ? cells : keys_out, vals_in, perm, n, 0, end_bit, s);
    if (e != cudaSuccess) return e;
    if (recs) {   // one random 32-byte read per point
        GatherDst dst{};
        for (int d = 0; d < P.n_in; ++d) dst.p[d] = t_sorted[d];
        switch (P.n_in) {
#define DIND_GR(NN) case NN: plan_gather_rec_kernel<T, NN><<<nb, BLOCK, 0, s>>>(recs, perm, dst, n); break;
            DIND_GR(1) DIND_GR(2) DIND_GR(3) DIND_GR(4) DIND_GR(5) DIND_GR(6) DIND_GR(7) DIND_GR(8)
#undef DIND_GR
            default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    for (int d = 0; d < P.n_in; ++d) {
        plan_gather_kernel<T><<<nb, BLOCK, 0, s>>>(P.dim[d].t_eval, perm, t_sorted[d], n);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template cudaError_t plan_sort_points<float>(const EvalParams<float>&, uint32_t*, float* const*, uint32_t*, int, void*, size_t*, float*, cudaStream_t);
template cudaError_t plan_sort_points<double>(const EvalParams<double>&, uint32_t*, double* const*, uint32_t*, int, void*, size_t*, double*, cudaStream_t);

}  // namespace dind
