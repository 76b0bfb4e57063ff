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

#include "dind_kernels.h"

namespace dind {

namespace {

constexpr int BLOCK = 256;

template <typename T>
__global__ void __launch_bounds__(BLOCK) plan_key_kernel(const __grid_constant__ EvalParams<T> P, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k >= P.n_points) return;
    int64_t base = 0;
    for (int d = 0; d < P.n_in; ++d) {
        const DimParams<T>& D = P.dim[d];
        base += (int64_t)dim_first_node<T>(D, __ldcs(D.t_eval + k)) * D.u_stride;
    }
    keys[k] = (uint32_t)base;
    vals[k] = (uint32_t)k;
}

template <typename T>
__global__ void __launch_bounds__(BLOCK) plan_gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ perm,
                                                            T* __restrict__ dst, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k < n) dst[k] = src[perm[k]];
}

__global__ void __launch_bounds__(BLOCK) plan_inverse_kernel(const uint32_t* __restrict__ perm, uint32_t* __restrict__ inv, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (k < n) inv[perm[k]] = (uint32_t)k;
}

// out[i + cc * n] = staged[inv[i] * ct + cc]: one random (ct * sizeof(T))-byte read per point, coalesced
// writes — instead of ct random read-modify-write sector accesses per point for a direct scatter.
template <typename T>
__global__ void __launch_bounds__(BLOCK) plan_unpermute_kernel(const T* __restrict__ staged, const uint32_t* __restrict__ inv,
                                                               T* __restrict__ out, int64_t n, int ct) {
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const T* src = staged + (int64_t)__ldcs(inv + i) * ct;
    for (int cc = 0; cc < ct; ++cc) __stcs(out + i + (int64_t)cc * n, __ldcs(src + cc));
}

}  // namespace

cudaError_t plan_inverse(const uint32_t* perm, uint32_t* inv, int64_t n, cudaStream_t s) {
    plan_inverse_kernel<<<(unsigned)((n + BLOCK - 1) / BLOCK), BLOCK, 0, s>>>(perm, inv, n);
    return cudaGetLastError();
}

template <typename T>
cudaError_t plan_unpermute(const T* staged, const uint32_t* inv, T* out, int64_t n, int ct, cudaStream_t s) {
    plan_unpermute_kernel<T><<<(unsigned)((n + BLOCK - 1) / BLOCK), BLOCK, 0, s>>>(staged, inv, out, n, ct);
    return cudaGetLastError();
}
template cudaError_t plan_unpermute<float>(const float*, const uint32_t*, float*, int64_t, int, cudaStream_t);
template cudaError_t plan_unpermute<double>(const double*, const uint32_t*, double*, int64_t, int, cudaStream_t);

template <typename T>
cudaError_t plan_sort_points(const EvalParams<T>& P, uint32_t* perm, T* const* t_sorted, void* scratch,
                             size_t* scratch_bytes, cudaStream_t s) {
    const int64_t n = P.n_points;
    int end_bit = 1;
    while (end_bit < 32 && (1LL << end_bit) < P.u_comp_stride) ++end_bit;
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
    plan_key_kernel<T><<<nb, BLOCK, 0, s>>>(P, keys_in, vals_in);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, vals_in, perm, n, 0, end_bit, s);
    if (e != cudaSuccess) return e;
    for (int d = 0; d < P.n_in; ++d) {
        plan_gather_kernel<T><<<nb, BLOCK, 0, s>>>(P.dim[d].t_eval, perm, t_sorted[d], n);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template cudaError_t plan_sort_points<float>(const EvalParams<float>&, uint32_t*, float* const*, void*, size_t*, cudaStream_t);
template cudaError_t plan_sort_points<double>(const EvalParams<double>&, uint32_t*, double* const*, void*, size_t*, cudaStream_t);

}  // namespace dind
