// FFMA against FFMA2 (fma.rn.f32x2, sm_100a): issue rate and flops of a full machine, 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b)
{
    if (PACKED) {
        unsigned long long acc[8], A, Bv;
        float2 av = make_float2(a, a), bv = make_float2(b, b);
        A = *reinterpret_cast<unsigned long long*>(&av);
        Bv = *reinterpret_cast<unsigned long long*>(&bv);
        for (int j = 0; j < 8; ++j) {
            float2 s = make_float2(threadIdx.x + j, threadIdx.x - j);
            acc[j] = *reinterpret_cast<unsigned long long*>(&s);
        }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(A), "l"(Bv));
        }
        float r = 0.f;
        for (int j = 0; j < 8; ++j) {
            float2 s = *reinterpret_cast<float2*>(&acc[j]);
            r += s.x + s.y;
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {
        float acc[8];
        for (int j = 0; j < 8; ++j) acc[j] = threadIdx.x + j;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[j]) : "f"(a), "f"(b));
        }
        float r = 0.f;
        for (int j = 0; j < 8; ++j) r += acc[j];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}

int main()
{
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (packed)
                k<true><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            else
                k<false><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double instr = (double)148 * 8 * 256 * iters * 8;
            printf("%s: %.3f ms, %.2f T thread-instructions/s, %.2f TFLOP/s\n", packed ? "FFMA2" : "FFMA ", ms,
                   instr / ms / 1e9, instr * (packed ? 4 : 2) / ms / 1e9);
        }
    }
    return 0;
}
