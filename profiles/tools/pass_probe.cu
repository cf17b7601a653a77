// pass_probe.cu — cycles per solver pass of one warp in isolation (tuning aid, not part of the library).
// Measured on B200 (one warp per scheduler): 136 cycles per pass with the three shuffles, 89 with a constant
// partner; a variant in which each lane carried its partner's twist forward itself (no shuffle in the sweep,
// ten more FP instructions) took 126 and lost 9 % in packed launches: the pass is bound by its instruction
// count, not by the exchange.  Two warps per scheduler: 176 / 120.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I manipulation_planning_suite_b200/csrc
//        -o gpurun_out/pass_probe profiles/tools/pass_probe.cu
#include "rollout.cu"
#include <cstdio>

template <int MODE>
__global__ void probe(float* out, long long* cyc, int passes, int partner_xor)
{
    const Ln ln = make_ln<32>();
    const int lane = ln.g;
    Body bd = {};
    bd.vx = 0.01f * lane;
    bd.vy = 0.02f;
    bd.w = 0.1f;
    bd.inv_m = 2.8f;
    bd.inv_i = 500.0f;
    MRows R;
    R.m0 = make_float4(0.f, 0.6f, 0.8f, 0.3f);
    R.r0 = make_float4(0.01f, -0.02f, 0.03f, 0.01f);
    R.g0 = make_float4(0.1f, 0.12f, 0.001f, 0.05f);
    R.r1 = make_float4(-0.01f, 0.02f, 0.03f, -0.01f);
    R.g1 = make_float4(0.1f, 0.12f, 0.001f, 0.05f);
    R.kc = make_float4(0.01f, 0.02f, 0.01f, 0.02f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int j = lane ^ partner_xor;
    const bool is_a = lane < j;
    __syncwarp();
    float ovx = 0.f, ovy = 0.f, ow_ = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < passes; ++i) {
        if (MODE == 0) {
            solve_core<true>(R, acc, bd, ln, true, j, false, is_a);
        } else if (MODE == 1) {
            // no exchange: the partner's twist is a constant (pure arithmetic chain)
            solve_core<true>(R, acc, bd, ln, true, lane, true, true);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = bd.vx + bd.vy + bd.w + acc.x + ovx + ovy + ow_;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    float* out;
    long long* cyc;
    cudaMalloc(&out, 4 * 32 * 1024);
    cudaMalloc(&cyc, 8 * 1024);
    const int passes = 4096;
    long long h[4];
    for (int warps = 1; warps <= 8; warps *= 2) {
        probe<0><<<1, 32 * warps>>>(out, cyc, passes, 1);
        probe<0><<<1, 32 * warps>>>(out, cyc, passes, 1);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double a = (double)h[0] / passes;
        probe<1><<<1, 32 * warps>>>(out, cyc, passes, 1);
        probe<1><<<1, 32 * warps>>>(out, cyc, passes, 1);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        double b = (double)h[0] / passes;
        printf("warps/CTA %d: cycles per pass with shuffles %.1f, constant partner %.1f\n", warps, a, b);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
