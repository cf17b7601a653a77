// icache_probe.cu — how much straight-line code can one warp loop over before instruction fetch shows?
// (tuning aid).  A loop body of N independent FFMA-class instructions is run by one warp per scheduler;
// cycles per instruction vs N gives the capacity of the per-scheduler instruction buffer and the cost of missing it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O1 -o profiles/tools/_bin/icache_probe profiles/tools/icache_probe.cu
#include <cstdio>

#define I8(a)                                   \
    asm volatile("fma.rn.f32 %0, %0, %8, %9;\n" \
                 "fma.rn.f32 %1, %1, %8, %9;\n" \
                 "fma.rn.f32 %2, %2, %8, %9;\n" \
                 "fma.rn.f32 %3, %3, %8, %9;\n" \
                 "fma.rn.f32 %4, %4, %8, %9;\n" \
                 "fma.rn.f32 %5, %5, %8, %9;\n" \
                 "fma.rn.f32 %6, %6, %8, %9;\n" \
                 "fma.rn.f32 %7, %7, %8, %9;\n" \
                 : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]) \
                 : "f"(m), "f"(c));
#define I64(a) I8(a) I8(a) I8(a) I8(a) I8(a) I8(a) I8(a) I8(a)
#define I256(a) I64(a) I64(a) I64(a) I64(a)
#define I1024(a) I256(a) I256(a) I256(a) I256(a)

template <int N64>
__global__ void probe(float* out, long long* cyc, int reps, float m, float c)
{
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    long long t0 = 0;
#pragma unroll 1
    for (int r = 0; r < reps + 1; ++r) {
        if (r == 1) t0 = clock64();  // first trip warms the caches
        if (N64 >= 2) { I64(a) I64(a) }
        if (N64 >= 4) { I64(a) I64(a) }
        if (N64 >= 6) { I64(a) I64(a) }
        if (N64 >= 8) { I64(a) I64(a) }
        if (N64 >= 12) { I256(a) }
        if (N64 >= 16) { I256(a) }
        if (N64 >= 24) { I256(a) I256(a) }
        if (N64 >= 32) { I256(a) I256(a) }
        if (N64 >= 48) { I1024(a) }
        if (N64 >= 64) { I1024(a) }
        if (N64 >= 96) { I1024(a) I1024(a) }
        if (N64 >= 128) { I1024(a) I1024(a) }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int N64>
void run(float* out, long long* cyc, int warps, int blocks)
{
    const int reps = 64;
    probe<N64><<<blocks, 32 * warps>>>(out, cyc, reps, 0.999f, 0.001f);
    probe<N64><<<blocks, 32 * warps>>>(out, cyc, reps, 0.999f, 0.001f);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  %5d instructions (%6.1f KB): %.2f cycles / instruction\n", N64 * 64, N64 * 64 * 16 / 1024.0,
           (double)h / reps / (N64 * 64));
}

int main()
{
    float* out;
    long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 1024);
    cudaMalloc(&cyc, 8 * 4096);
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int warps = cfg == 0 ? 1 : 4, blocks = cfg == 2 ? 148 * 3 : 1;
        printf("%d warp(s) per CTA, %d CTA(s):\n", warps, blocks);
        run<2>(out, cyc, warps, blocks);
        run<4>(out, cyc, warps, blocks);
        run<6>(out, cyc, warps, blocks);
        run<8>(out, cyc, warps, blocks);
        run<12>(out, cyc, warps, blocks);
        run<16>(out, cyc, warps, blocks);
        run<24>(out, cyc, warps, blocks);
        run<32>(out, cyc, warps, blocks);
        run<48>(out, cyc, warps, blocks);
        run<64>(out, cyc, warps, blocks);
        run<96>(out, cyc, warps, blocks);
        run<128>(out, cyc, warps, blocks);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
