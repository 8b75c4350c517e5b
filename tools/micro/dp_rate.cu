// Micro-benchmark: issue rate of dependent-free fp64 / fp32 FMA streams on one GPU (development aid).
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void fma_kernel(T* out, int iters) {
    T a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T b = (T)1.0000001, c = (T)0.5;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <typename T> double run(const char* name) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, threads = 512, iters = 20000;
    T* out; cudaMalloc(&out, sizeof(T) * blocks * threads);
    fma_kernel<T><<<blocks, threads>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); fma_kernel<T><<<blocks, threads>>>(out, iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * threads * iters * 8;
    printf("%s: %.3f ms, %.2f TFMA/s (%.2f TFLOP/s), %.1f FMA/clk/SM at 1.965 GHz\n", name, ms, fmas / ms / 1e9,
           2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out); return ms;
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
