// DFMA throughput of one B200 (sm_100a): the denominator for "FP64 pipe utilisation" statements
// (no fp64 entry in MEASURED_PEAKS.json).  Each thread runs 8 independent fma chains.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/fp64_peak scripts/micro/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, blocks = sms * 2, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 8 * (double)iters * threads * blocks;
    printf("{\"fp64_dfma_tflops\": %.3f, \"sms\": %d, \"ms\": %.4f, \"err\": \"%s\"}\n", flops / (best * 1e-3) / 1e12, sms, best,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
