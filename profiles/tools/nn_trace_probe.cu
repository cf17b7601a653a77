// nn_trace_probe.cu - where an isolated NN scan launch spends its time (B = 11, 1 M nodes): per-CTA %globaltimer
// stamps of the kernel's phases (NN_TRACE build of nn_scan.cu) against the CUDA-event duration of the same launch.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNN_TRACE -o nn_trace_probe nn_trace_probe.cu
#include "../../manipulation_planning_suite_b200/csrc/nn_scan.cu"

#include <algorithm>
#include <cstdio>
#include <vector>

__global__ void empty_kernel(int* p)
{
    if (p && threadIdx.x == 1000) *p = 1;
}

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (x);                                                                 \
        if (e__ != cudaSuccess) {                                                              \
            printf("CUDA error %s at %d: %s\n", cudaGetErrorString(e__), __LINE__, #x);        \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

int main(int argc, char** argv)
{
    const int B = 11;
    const long long N = argc > 1 ? atoll(argv[1]) : 1000000;
    const size_t cap = (size_t)((N + 1023) / 1024 * 1024);
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int NT = 3;
    TreeView tv[NT];
    long long* d_size;
    CK(cudaMalloc(&d_size, 8));
    for (int t = 0; t < NT; ++t) {
        tv[t].B = B;
        tv[t].cap = cap;
        CK(cudaMalloc(&tv[t].states, sizeof(float) * 3 * B * cap));
        CK(cudaMalloc(&tv[t].controls, sizeof(float) * 5 * cap));
        CK(cudaMalloc(&tv[t].parent, 4 * cap));
        CK(cudaMalloc(&tv[t].slice, 4 * cap));
        CK(mps_launch_tree_fill(tv[t], N, 1003 + t, 64, -1.5f, 1.5f, -1.5f, 1.5f, d_size, st));
    }
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NNParams p;
    memset(&p, 0, sizeof(p));
    p.n = N;
    p.pw = 1.0;
    p.ow = 0.08;
    p.Q = 1;
    p.grid_x = mps_nn_grid_x(N, sms);
    CK(cudaMalloc(&p.out_idx, 64));
    CK(cudaMalloc(&p.out_dist, 64));
    CK(cudaMalloc(&p.part_d, 8 * 4096));
    CK(cudaMalloc(&p.part_i, 8 * 4096));
    CK(cudaMalloc(&p.tickets, 64));
    CK(cudaMemset(p.tickets, 0, 64));
    p.inline_q = 1;
    p.inline_mask = (1u << B) - 1u;
    p.inline_filter = -1;
    for (int i = 0; i < 3 * B; ++i) p.q_inline[i] = 0.1f * (float)((i * 7) % 11) - 0.5f;
    p.abs_eps = 1e-6f;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 6; ++w) {
        p.tree = tv[w % NT];
        CK(mps_launch_nn_scan(p, st));
    }
    CK(cudaStreamSynchronize(st));
    const int G = p.grid_x;
    std::vector<unsigned long long> tr(8 * 4096);
    double acc[8] = {0};
    double ev_acc = 0, pro[2] = {0, 0};
    const int reps = 20;
    for (int r = 0; r < reps; ++r) {
        p.tree = tv[r % NT];
        CK(cudaEventRecord(e0, st));
        CK(mps_launch_nn_scan(p, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpyFromSymbol(tr.data(), g_nn_trace, sizeof(unsigned long long) * 8 * G));
        unsigned long long t7max = 0, t2wmax = 0;
        for (int c = 0; c < G; ++c) {
            t7max = std::max(t7max, tr[8 * c + 7]);
            t2wmax = std::max(t2wmax, tr[8 * c + 2]);
        }
        unsigned long long t0min = ~0ull, t0max = 0, t1max = 0, t2min = ~0ull, t2max = 0, t4max = 0, t5max = 0, t6 = 0;
        std::vector<unsigned long long> t2s;
        for (int c = 0; c < G; ++c) {
            t0min = std::min(t0min, tr[8 * c + 0]);
            t0max = std::max(t0max, tr[8 * c + 0]);
            t1max = std::max(t1max, tr[8 * c + 1]);
            t2min = std::min(t2min, tr[8 * c + 3]);
            t2max = std::max(t2max, tr[8 * c + 3]);
            t4max = std::max(t4max, tr[8 * c + 4]);
            t5max = std::max(t5max, tr[8 * c + 5]);
            t6 = std::max(t6, tr[8 * c + 6]);
            t2s.push_back(tr[8 * c + 3]);
        }
        std::sort(t2s.begin(), t2s.end());
        const double v[8] = {(double)(t0max - t0min), (double)(t1max - t0min), (double)(t2min - t0min),
                             (double)(t2s[G / 2] - t0min), (double)(t2max - t0min), (double)(t4max - t0min),
                             (double)(t5max - t0min), (double)(t6 - t0min)};
        if (r >= 2) {
            pro[0] += (double)(t2wmax - t0min);
            pro[1] += (double)(t7max - t0min);
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
            ev_acc += ms * 1e6;
        }
    }
    const int n = reps - 2;
    printf("N=%lld grid=%d: event pair %.2f us; inside the kernel, ns after the first CTA's first instruction: last CTA starts %.0f, "
           "last prologue done %.0f, first / median / last CTA done streaming %.0f / %.0f / %.0f, last deferred pass + CTA "
           "reduction done %.0f, last hand-off done %.0f, result written %.0f\n",
           N, G, ev_acc / n / 1e3, acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n, acc[4] / n, acc[5] / n, acc[6] / n,
           acc[7] / n);
    printf("  prologue: dependency wait passed %.0f, prefetches issued %.0f\n", pro[0] / n, pro[1] / n);
    // the same with an empty kernel of the same shape between the events: what a launch costs whatever it does
    double em = 0;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0, st));
        empty_kernel<<<G, NN_THREADS, 0, st>>>(nullptr);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) em += ms * 1e3;
    }
    printf("  empty kernel, same grid, one event pair per launch: %.2f us\n", em / n);
    return 0;
}
