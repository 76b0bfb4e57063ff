// Microbenchmark (not part of the product): what does a random gather of aligned 512-byte records cost on this
// B200, as a function of how the lanes of a warp share a record and of the table size?
//   L = lanes per record (32 / L records per warp instruction), each lane loads 512 / L bytes as 32- or 16-byte loads.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/microbench_gather.cu -o scripts/microbench_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float4 ld16(const void* p) {
    float4 q;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p));
    return q;
}
__device__ __forceinline__ float ld32sum(const void* p) {
    float q[8];
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(q[0]), "=f"(q[1]), "=f"(q[2]), "=f"(q[3]), "=f"(q[4]), "=f"(q[5]), "=f"(q[6]), "=f"(q[7]) : "l"(p));
    return ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
}

// L lanes per record; V = bytes per load (16 or 32); each lane loads 512 / L bytes, the loads of one instruction
// cover V * L contiguous bytes of the record
template <int L, int V>
__global__ void __launch_bounds__(256) gather(const char* __restrict__ tab, uint32_t n_rec, float* __restrict__ out, int64_t n_pts, uint32_t seed) {
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t pt = t / L;
    const int lane = (int)(t % L);
    if (pt >= n_pts) return;
    const uint32_t r = (uint32_t)(((uint64_t)hash32((uint32_t)pt * 2654435761u + seed) * n_rec) >> 32);
    const char* p = tab + (size_t)r * 512 + lane * V;
    float acc = 0.f;
    constexpr int NL = 512 / (L * V);
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        if (V == 32) acc += ld32sum(p + j * (L * V));
        else { const float4 q = ld16(p + j * (L * V)); acc += (q.x + q.y) + (q.z + q.w); }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) __stcs(out + pt, acc);
}

// the config-5 pattern of the component-fastest copy: 16 row segments of 32 bytes in 16 distant places
__global__ void __launch_bounds__(256) gather_rows(const char* __restrict__ tab, float* __restrict__ out, int64_t n_pts, uint32_t seed) {
    const int64_t pt = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (pt >= n_pts) return;
    uint32_t h = hash32((uint32_t)pt * 2654435761u + seed);
    int i[5];
    for (int d = 0; d < 5; ++d) { i[d] = (int)(h % 31u); h = hash32(h + d); }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const size_t node = i[0] + 32ull * ((i[1] + (c & 1)) + 32ull * ((i[2] + ((c >> 1) & 1)) + 32ull * ((i[3] + ((c >> 2) & 1)) + 32ull * (i[4] + (c >> 3)))));
        const float4 a = ld16(tab + node * 16), b = ld16(tab + node * 16 + 16);
        acc += (a.x + a.y) + (b.z + b.w);
    }
    __stcs(out + pt, acc);
}

int main() {
    const int64_t n_pts = 100000000;
    float* out;
    CK(cudaMalloc(&out, n_pts * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (double gb : {0.0625, 0.125, 0.5, 2.0, 8.0, 32.0}) {
        const size_t bytes = (size_t)(gb * (1ull << 30));
        char* tab;
        CK(cudaMalloc(&tab, bytes));
        CK(cudaMemset(tab, 0, bytes));
        const uint32_t n_rec = (uint32_t)(bytes / 512);
        auto time = [&](const char* name, auto fn) {
            fn(1u);
            cudaEventRecord(e0);
            const int K = 3;
            for (int i = 0; i < K; ++i) fn(7u + i);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t e = cudaGetLastError();
            printf("table %6.3f GB  %-26s %8.3f ms  %6.2f Grec/s  %7.1f GB/s %s\n", gb, name, ms / K, n_pts / (ms / K * 1e-3) / 1e9,
                   n_pts * 512.0 / (ms / K * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
#define RUN(L, V) time("lanes/rec " #L ", " #V " B loads", [&](uint32_t s) { gather<L, V><<<(unsigned)((n_pts * L + 255) / 256), 256>>>(tab, n_rec, out, n_pts, s); });
        RUN(1, 32) RUN(1, 16) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(16, 32) RUN(16, 16) RUN(32, 16)
        if (gb == 0.5) time("16 row segments (aos)", [&](uint32_t s) { gather_rows<<<(unsigned)((n_pts + 255) / 256), 256>>>(tab, out, n_pts, s); });
        cudaFree(tab);
    }
    return 0;
}
