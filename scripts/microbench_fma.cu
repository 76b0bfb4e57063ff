// Microbenchmark (not part of the product): register-resident FP32 / FP64 FMA peak on this B200,
// the denominator for the FMA-bound figures (config 3 / config 4) — BASELINE.md §2.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_peak(T* out, int iters) {
    T a0 = threadIdx.x * T(1e-3), a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = T(1.0000001), c = T(1e-7);
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T>
void run(const char* name) {
    T* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(T));
    const int iters = 1 << 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_peak<T><<<148 * 8, 256>>>(out, iters);
    cudaEventRecord(e0);
    fma_peak<T><<<148 * 8, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = 148.0 * 8 * 256 * 8 * iters;
    printf("%s FMA peak: %.2f TFMA/s = %.2f TFLOP/s (%.2f ms)\n", name, fmas / (ms * 1e-3) / 1e12, 2 * fmas / (ms * 1e-3) / 1e12, ms);
    cudaFree(out);
}
int main() {
    run<float>("FP32");
    run<double>("FP64");
    return 0;
}
