// emu_kernels.cpp — the repo's CUDA kernels compiled for the host through the SIMT emulator
// (TEST INFRASTRUCTURE, see cuda_emu.h).  The kernel sources are included as they are; only the launch
// glue below is specific to this file.  Entry points mirror the C ABI's structs so that tests can run the
// same inputs through the oracle, through this build of the kernel source and (on a GPU box) through
// libmps_b200.so.
#define MPS_CPU_EMU 1
#include "cuda_emu.h"

#include "../../manipulation_planning_suite_b200/csrc/scene_derive.h"
#include "../../manipulation_planning_suite_b200/csrc/rollout.cu"

#include <vector>

namespace {

struct Buffers {
    std::vector<float> states, rest;
    std::vector<uint32_t> flags;
    std::vector<int32_t> steps, nok, goal_step;
    std::vector<double> dist;
    std::vector<uint8_t> redo;
    std::vector<unsigned char> record;
    unsigned int work = 0;
    unsigned long long counters[8] = {0};
};

template <int WARPS, class Body>
void run_grid(RolloutParams& p, int groups, int ctas, Body body)
{
    const size_t smem = p.scene_bytes + (size_t)groups * p.warp_bytes;
    *p.work_counter = 0u;
    int want = (p.K + groups - 1) / groups;
    int grid = ctas < want ? ctas : want;
    if (grid < 1) grid = 1;
    emu::launch(dim3(grid), dim3(WARPS * 32), smem, body);
}

template <int WARPS, int MINB, int G, bool ROLLED = false>
void run_rollout(RolloutParams p, int ctas)
{
    if (G < 32 && p.mcap_small > 0) p.warp_bytes = p.warp_bytes_small;
    run_grid<WARPS>(p, WARPS * (32 / G), ctas, [&]() { rollout_kernel<WARPS, MINB, G, ROLLED>(p); });
}

}  // namespace

extern "C" {

// K chains + argmin, as mps_extend (abi.cu) drives rollout_kernel and select_kernel.
//   group: lanes per chain (32, 16, 8); minb: register-budget instantiation (1 | 3 | 4, 5 = 4 with rolled sweeps) of the
//   one-chain-per-warp build; mcap_small: manifold slots of a packed launch (0 = the scene's); warps: warps per emulated CTA;
//   ctas: CTAs of the emulated grid (chains are pulled from the work counter as on the device)
int emu_extend(const mps_scene* sc, const mps_extend_in* in, const mps_extend_out* out, int group, int minb,
               int mcap_small, int ctas, uint64_t* counters_out)
{
    DevScene ds;
    derive_scene(sc, &ds);
    const int B = sc->n_bodies, K = in->K, L = in->L;
    const size_t KL = (size_t)K * L;
    Buffers b;
    b.states.assign(KL * 3 * B, -7.0f);
    b.rest.assign(KL, -7.0f);
    b.flags.assign(KL, 0xdeadbeefu);
    b.steps.assign(KL, -7);
    b.nok.assign(K, -7);
    b.goal_step.assign(K, -7);
    b.dist.assign(K, -7.0);
    b.redo.assign(K, 0);
    const size_t rec_bytes = sizeof(mps_extend_best) + sizeof(float) * L * (3 * B + MPS_CONTROL_DIM) + sizeof(uint32_t) * L + 16;
    b.record.assign(rec_bytes, 0);

    RolloutParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.scene = &ds;
    rp.scene_bytes = mps_rollout_scene_bytes();
    rp.warp_bytes = mps_rollout_warp_bytes(ds);
    if (mcap_small >= ds.Mcap) mcap_small = 0;
    rp.mcap_small = mcap_small;
    rp.warp_bytes_small = mcap_small ? mps_rollout_group_bytes(ds, mcap_small) : rp.warp_bytes;
    rp.redo = b.redo.data();
    rp.start = in->starts ? in->starts : in->start;
    rp.start_stride = in->starts ? (size_t)3 * B : 0;
    rp.stop_at_goal = in->stop_at_goal;
    rp.controls = in->controls;
    rp.n_ctrl = in->n_ctrl;
    rp.K = K;
    rp.L = L;
    rp.dest = in->dest;
    rp.weights = in->weights;
    rp.mask = in->mask;
    rp.goals = in->goals;
    rp.n_goals = in->n_goals;
    rp.ball_center = in->ball_center;
    rp.ball_radius = in->ball_radius;
    rp.out_states = b.states.data();
    rp.out_rest = b.rest.data();
    rp.out_flags = b.flags.data();
    rp.out_steps = b.steps.data();
    rp.out_dist = b.dist.data();
    rp.out_nok = b.nok.data();
    rp.out_goal_step = b.goal_step.data();
    rp.work_counter = &b.work;
    rp.counters = b.counters;
    rp.out_cycles = nullptr;
    if (ctas < 1) ctas = 2;

    if (group == 32) {
        rp.mcap_small = 0;
        rp.redo_only = 0;
        if (minb == 1)
            run_rollout<2, 1, 32>(rp, ctas);
        else if (minb == 5)  // the 128-register build with the sweeps rolled over the rounds
            run_rollout<2, 4, 32, true>(rp, ctas);
        else if (minb == 3)
            run_rollout<2, 3, 32>(rp, ctas);
        else
            run_rollout<2, 4, 32>(rp, ctas);
    } else {
        RolloutParams q = rp;
        q.redo_only = 0;
        if (group == 8)
            run_rollout<2, 4, 8>(q, ctas);
        else
            run_rollout<2, 4, 16, true>(q, ctas);  // two chains per warp: always the rolled sweeps
        if (rp.mcap_small > 0) {
            rp.redo_only = 1;
            rp.mcap_small = 0;
            run_rollout<2, 4, 32>(rp, ctas);
        }
    }

    SelectParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.K = K;
    sp.L = L;
    sp.B = B;
    sp.M = 1;
    sp.has_dest = in->dest ? 1 : 0;
    sp.compare_f32 = in->compare_f32;
    sp.k_offset = in->k_offset;
    sp.dist = rp.out_dist;
    sp.nok = rp.out_nok;
    sp.goal_step = rp.out_goal_step;
    sp.states = rp.out_states;
    sp.rest = rp.out_rest;
    sp.flags = rp.out_flags;
    sp.controls = rp.controls;
    unsigned char* rec = b.record.data();
    sp.best = (mps_extend_best*)rec;
    sp.best_states = (float*)(rec + sizeof(mps_extend_best));
    sp.best_controls = sp.best_states + (size_t)L * 3 * B;
    sp.best_flags = (uint32_t*)(sp.best_controls + (size_t)L * MPS_CONTROL_DIM);
    sp.append = 0;
    int threads = K >= 1024 ? 1024 : ((K + 31) / 32) * 32;
    if (threads < 32) threads = 32;
    emu::launch(dim3(1), dim3(threads), 0, [&]() { select_kernel(sp); });

    if (out->best) memcpy(out->best, sp.best, sizeof(mps_extend_best));
    if (out->best_states) memcpy(out->best_states, sp.best_states, sizeof(float) * 3 * B * L);
    if (out->best_controls) memcpy(out->best_controls, sp.best_controls, sizeof(float) * MPS_CONTROL_DIM * L);
    if (out->best_flags) memcpy(out->best_flags, sp.best_flags, sizeof(uint32_t) * L);
    if (out->all_states) memcpy(out->all_states, b.states.data(), sizeof(float) * b.states.size());
    if (out->all_rest) memcpy(out->all_rest, b.rest.data(), sizeof(float) * KL);
    if (out->all_flags) memcpy(out->all_flags, b.flags.data(), sizeof(uint32_t) * KL);
    if (out->all_dist) memcpy(out->all_dist, b.dist.data(), sizeof(double) * K);
    if (out->all_steps) memcpy(out->all_steps, b.steps.data(), sizeof(int32_t) * KL);
    if (counters_out) memcpy(counters_out, b.counters, sizeof(b.counters));
    return 0;
}

// isValid for n states (validity_kernel)
int emu_is_valid_batch(const mps_scene* sc, const float* states, int n, uint8_t* valid, uint32_t* flags)
{
    DevScene ds;
    derive_scene(sc, &ds);
    ValidityParams vp;
    memset(&vp, 0, sizeof(vp));
    vp.scene = &ds;
    vp.scene_bytes = mps_rollout_scene_bytes();
    vp.warp_bytes = mps_rollout_warp_bytes(ds);
    vp.states = states;
    vp.n = n;
    std::vector<uint8_t> v(n);
    std::vector<uint32_t> f(n);
    vp.valid = v.data();
    vp.flags = f.data();
    const int W = 2;
    int grid = (n + W - 1) / W;
    if (grid > 2) grid = 2;
    emu::launch(dim3(grid), dim3(W * 32), vp.scene_bytes + W * vp.warp_bytes, [&]() { validity_kernel<W>(vp); });
    if (valid) memcpy(valid, v.data(), n);
    if (flags) memcpy(flags, f.data(), sizeof(uint32_t) * n);
    return 0;
}

}  // extern "C"
