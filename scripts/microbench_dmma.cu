// Microbenchmark (not part of the product): FP64 tensor-core MMA (mma.sync.m8n8k4.f64, SASS DMMA) rate on this
// B200, alone and interleaved with DFMA — decides whether stage 1 of the cell-sorted cubic B-spline kernel
// (a (points x 4) x (4 x 48) product per cell) belongs on the tensor pipe.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: DMMA only, 1: DFMA only, 2: both interleaved (same instruction counts as 0 + 1)
__global__ void k(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    double c[16];
    double f[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 1) dmma(c[2 * i], c[2 * i + 1], a, b);
            if (MODE != 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fma(f[j], b, a);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
    double* out;
    const int blocks = 148 * 4, threads = 256, iters = 1 << 13;
    cudaMalloc(&out, blocks * threads * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 16);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double n_dmma = MODE != 1 ? warps * iters * 8 : 0;               // warp-level DMMA instructions, 256 FMA each
    const double n_dfma = MODE != 0 ? warps * iters * 8 * 8 * 32.0 : 0;    // lane-level DFMA
    printf("%-28s %8.3f ms  DMMA %.2f TFMA/s (%.2f clk/SM per DMMA)  DFMA %.2f TFMA/s\n", name, ms,
           n_dmma * 256 / (ms * 1e-3) / 1e12, n_dmma ? ms * 1e-3 * 1.965e9 * 148 / n_dmma : 0.0, n_dfma / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("DMMA m8n8k4 only");
    run<1>("DFMA only (64 per DMMA slot)");
    run<2>("DMMA + DFMA interleaved");
    return 0;
}
