// Microbenchmark (not part of the product): what do the output-store patterns of the grid kernels
// cost by themselves on this B200?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void fill_linear(float4* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) __stcs(out + i, make_float4(1.f, 2.f, 3.f, 4.f));
}
__global__ void fill_linear_plain(float4* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
// pencil pattern: thread rv writes float4 at rv*4 + g*R for g in chunk
template <bool CS>
__global__ void fill_pencil(float* out, int RV, long R, int chunk) {
    int rv = blockIdx.x * blockDim.x + threadIdx.x;
    if (rv >= RV) return;
    float* op = out + (long)rv * 4 + (long)blockIdx.y * chunk * R;
    for (int g = 0; g < chunk; ++g, op += R) {
        float4 v = make_float4(g, rv, 1.f, 2.f);
        if (CS) __stcs((float4*)op, v); else *(float4*)op = v;
    }
}
// pencil-pattern stores + L2-resident reads of a 67 MB array (about the read volume of the real kernel)
__global__ void fill_pencil_reads(float* out, const float* __restrict__ u, int RV, long R, int chunk) {
    int rv = blockIdx.x * blockDim.x + threadIdx.x;
    if (rv >= RV) return;
    float* op = out + (long)rv * 4 + (long)blockIdx.y * chunk * R;
    // u index: row of 256 floats per (j, k) pair, like a 2x-upsampled trilinear read
    const int i = (rv % 128) * 2, j = (rv / 128) / 2;
    for (int g = 0; g < chunk; ++g, op += R) {
        const int k = (blockIdx.y * chunk + g) / 2;
        const float* p = u + ((long)k * 256 + j) * 256 + i;
        float a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        __stcs((float4*)op, make_float4(a, a + b, b, b + c));
    }
}
// persistent grid-stride copy for reference
__global__ void copy_k(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
int main() {
    const size_t n = 512ull * 512 * 512;
    float *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 0, n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto fn, double bytes) {
        for (int i = 0; i < 3; ++i) fn();
        cudaEventRecord(e0);
        const int K = 20;
        for (int i = 0; i < K; ++i) fn();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %8.1f us  %7.1f GB/s\n", name, ms / K * 1e3, bytes / (ms / K * 1e-3) / 1e9);
        return 0;
    };
    size_t n4 = n / 4;
    time("fill linear stcs", [&] { fill_linear<<<(unsigned)((n4 + 255) / 256), 256>>>((float4*)b, n4); }, n * 4.0);
    time("fill linear plain", [&] { fill_linear_plain<<<(unsigned)((n4 + 255) / 256), 256>>>((float4*)b, n4); }, n * 4.0);
    for (int chunk : {8, 32, 64, 512}) {
        int RV = 512 * 512 / 4; long R = 512 * 512;
        dim3 g((RV + 127) / 128, 512 / chunk);
        char nm[64];
        snprintf(nm, 64, "fill pencil stcs chunk=%d", chunk);
        time(nm, [&] { fill_pencil<true><<<g, 128>>>(b, RV, R, chunk); }, n * 4.0);
        snprintf(nm, 64, "fill pencil plain chunk=%d", chunk);
        time(nm, [&] { fill_pencil<false><<<g, 128>>>(b, RV, R, chunk); }, n * 4.0);
    }
    for (int chunk : {32, 64}) {
        int RV = 512 * 512 / 4; long R = 512 * 512;
        dim3 g((RV + 127) / 128, 512 / chunk);
        char nm[64];
        snprintf(nm, 64, "pencil stores + reads chunk=%d", chunk);
        time(nm, [&] { fill_pencil_reads<<<g, 128>>>(b, a, RV, R, chunk); }, n * 4.0 + 67108864.0);
        // same with residency limited by dynamic shared memory (like the real kernel's 5 CTAs = 20 warps / SM)
        for (int ctas : {12, 8, 5, 3}) {
            int smem = (227 * 1024) / ctas - 1024;
            cudaFuncSetAttribute(fill_pencil_reads, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            snprintf(nm, 64, "  ... %d CTAs/SM chunk=%d", ctas, chunk);
            time(nm, [&] { fill_pencil_reads<<<g, 128, smem>>>(b, a, RV, R, chunk); }, n * 4.0 + 67108864.0);
        }
    }
    time("copy grid-stride (r+w)", [&] { copy_k<<<148 * 8, 256>>>((const float4*)a, (float4*)b, n4); }, n * 8.0);
    time("cudaMemcpy D2D (r+w)", [&] { cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice); }, n * 8.0);
    time("cudaMemset", [&] { cudaMemsetAsync(b, 0, n * 4); }, n * 4.0);
    return 0;
}
