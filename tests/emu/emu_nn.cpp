// emu_nn.cpp — nn_topk_kernel / nn_next_kernel of csrc/nn_scan.cu compiled for the host through the SIMT
// emulator (TEST INFRASTRUCTURE, see cuda_emu.h): the exact k-nearest-neighbour scan can be checked against a
// sorted double-precision scan on a machine without a GPU.  Launch glue only; the kernel source is included as is.
#define MPS_CPU_EMU 1
#include "cuda_emu.h"

#include "../../manipulation_planning_suite_b200/csrc/nn_scan.cu"

#include <vector>

extern "C" {

// states: node-major [n][3B] (transposed here into the device layout); returns the number of results; *status = the
// kernel's own verdict (0 exact, 1 ambiguous -> the library would fall back to successor scans)
int emu_nearest_k(const mps_scene* sc, const float* states, const int32_t* slice_ids, long long n, const float* query,
                  const float* weights, uint32_t mask, int32_t slice_filter, int k, int grid_x,
                  int64_t* indices, double* distances, int* status)
{
    const int B = sc->n_bodies;
    const size_t cap = (size_t)((n + 1023) / 1024 * 1024 + 1024);
    std::vector<float> soa((size_t)3 * B * cap, 0.0f);
    std::vector<int32_t> slice(cap, -1);
    for (long long i = 0; i < n; ++i) {
        for (int r = 0; r < 3 * B; ++r) soa[(size_t)r * cap + i] = states[(size_t)i * 3 * B + r];
        if (slice_ids) slice[i] = slice_ids[i];
    }
    TopKParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.nn.tree.states = soa.data();
    tp.nn.tree.slice = slice.data();
    tp.nn.tree.cap = cap;
    tp.nn.tree.B = B;
    tp.nn.n = n;
    tp.nn.pw = sc->position_weight;
    tp.nn.ow = sc->orientation_weight;
    tp.nn.Q = 1;
    tp.nn.inline_q = 1;
    tp.nn.inline_has_w = weights ? 1 : 0;
    tp.nn.inline_mask = mask;
    tp.nn.inline_filter = slice_filter;
    memcpy(tp.nn.q_inline, query, sizeof(float) * 3 * B);
    if (weights) memcpy(tp.nn.w_inline, weights, sizeof(float) * B);
    tp.nn.abs_eps = 1.0e-15f + 1.5e-6f * (float)sc->orientation_weight * (float)B * 4.0f;
    tp.nn.grid_x = grid_x;
    unsigned int ticket = 0;
    tp.nn.tickets = &ticket;
    std::vector<unsigned long long> keys((size_t)grid_x * NN_TOPK_MAX);
    std::vector<unsigned int> rej(grid_x);
    std::vector<long long> oi(NN_TOPK_MAX, -1);
    std::vector<double> od(NN_TOPK_MAX, 0.0);
    int meta[2] = {0, 0};
    tp.k = k;
    tp.part_keys = keys.data();
    tp.part_rej = rej.data();
    tp.out_idx = oi.data();
    tp.out_dist = od.data();
    tp.out_meta = meta;
    emu::launch(dim3(grid_x), dim3(NN_THREADS), 0, [&]() { nn_topk_kernel<true>(tp); });
    for (int i = 0; i < meta[0]; ++i) {
        indices[i] = oi[i];
        distances[i] = od[i];
    }
    *status = meta[1];
    return meta[0];
}

// successor scan: smallest (distance, index) above (after_d, after_i)
long long emu_nearest_next(const mps_scene* sc, const float* states, const int32_t* slice_ids, long long n,
                           const float* query, const float* weights, uint32_t mask, int32_t slice_filter, double after_d,
                           long long after_i, int grid_x, double* distance)
{
    const int B = sc->n_bodies;
    const size_t cap = (size_t)((n + 1023) / 1024 * 1024 + 1024);
    std::vector<float> soa((size_t)3 * B * cap, 0.0f);
    std::vector<int32_t> slice(cap, -1);
    for (long long i = 0; i < n; ++i) {
        for (int r = 0; r < 3 * B; ++r) soa[(size_t)r * cap + i] = states[(size_t)i * 3 * B + r];
        if (slice_ids) slice[i] = slice_ids[i];
    }
    NextParams np;
    memset(&np, 0, sizeof(np));
    np.nn.tree.states = soa.data();
    np.nn.tree.slice = slice.data();
    np.nn.tree.cap = cap;
    np.nn.tree.B = B;
    np.nn.n = n;
    np.nn.pw = sc->position_weight;
    np.nn.ow = sc->orientation_weight;
    np.nn.inline_q = 1;
    np.nn.inline_has_w = weights ? 1 : 0;
    np.nn.inline_mask = mask;
    np.nn.inline_filter = slice_filter;
    memcpy(np.nn.q_inline, query, sizeof(float) * 3 * B);
    if (weights) memcpy(np.nn.w_inline, weights, sizeof(float) * B);
    np.nn.grid_x = grid_x;
    unsigned int ticket = 0;
    np.nn.tickets = &ticket;
    std::vector<double> pd(grid_x);
    std::vector<long long> pi(grid_x);
    long long oi = -1;
    double od = 0.0;
    np.nn.part_d = pd.data();
    np.nn.part_i = pi.data();
    np.nn.out_idx = &oi;
    np.nn.out_dist = &od;
    np.after_d = after_d;
    np.after_i = after_i;
    emu::launch(dim3(grid_x), dim3(NN_THREADS), 0, [&]() { nn_next_kernel(np); });
    *distance = od;
    return oi;
}

}  // extern "C"
