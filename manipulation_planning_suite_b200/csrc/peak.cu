// peak.cu — FP32 FMA microbenchmark: the measured denominator of the rollout kernel's
// roofline fraction (MEASURED_PEAKS.json holds HBM and bf16 peaks only).
#include <cuda_runtime.h>

#include "peak.h"

namespace {
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f;
    float x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

cudaError_t mps_run_fma_peak(int num_sms, cudaStream_t stream, float* tflops)
{
    const int blocks = num_sms * 8, threads = 256, iters = 1 << 14;
    float* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(float) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 0.0f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, stream);
        fma_peak_kernel<<<blocks, threads, 0, stream>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    *tflops = best;
    return cudaGetLastError();
}
