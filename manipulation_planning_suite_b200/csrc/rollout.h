// rollout.h — launch interface of rollout.cu (internal to libmps_b200.so)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mps_b200.h"

struct DevScene;

struct RolloutParams {
    const DevScene* scene;  // device copy
    size_t scene_bytes;     // shared-memory bytes reserved for the scene copy
    size_t warp_bytes;      // shared-memory bytes per chain (lane group) with the scene's manifold capacity
    size_t warp_bytes_small;  // the same with mcap_small slots
    int mcap_small;         // manifold slots per chain of a packed launch (0: the scene's capacity)
    int redo_only;          // set by the launcher: only the chains flagged in `redo` are run
    int active_warps;       // tuning aid (MPS_ROLLOUT_ACTIVE_WARPS): > 0 = only this many warps of a CTA pull chains
    uint8_t* redo;          // [K] chains a packed launch left to the one-chain-per-warp kernel
    const float* start;     // [3B], or [K][3B] when start_stride != 0
    size_t start_stride;    // 0 or 3B
    const float* controls;  // [K][L][5]
    const int32_t* n_ctrl;  // [K] or nullptr
    int K, L;
    int stop_at_goal;       // chains end at their first goal state
    const float* dest;     // [3B] or nullptr
    const float* weights;  // [B] or nullptr
    uint32_t mask;
    const mps_goal* goals;
    int n_goals;
    const float* ball_center;  // [3B] or nullptr
    float ball_radius;
    float* out_states;    // [K][L][3B]
    float* out_rest;      // [K][L]
    uint32_t* out_flags;  // [K][L]
    int32_t* out_steps;   // [K][L]
    double* out_dist;     // [K]
    int32_t* out_nok;     // [K]
    int32_t* out_goal_step;  // [K]
    unsigned int* work_counter;
    const int32_t* order;   // [K] order in which the chains are pulled (longest first) or nullptr (0, 1, 2, ...)
    unsigned long long* counters;  // [4] or nullptr
    long long* out_cycles;         // [K] SM clock cycles spent on each chain (profiling aid) or nullptr
    // multi-instance batches (BASELINE cfg 5): K = M * K_per chains, instance m = k / K_per starts from node
    // parents[m] of forest segment inst_ids[m] and measures its distance to dest + m * 3B under masks[m]
    int K_per;                // 0 = single instance
    const int32_t* inst_ids;  // [M]
    const int32_t* parents;   // [M]
    const float* f_states;    // forest states, dimension-major
    size_t f_cols, f_seg;     // row stride of the forest / columns per segment
    const uint32_t* masks;    // [M] or nullptr (= mask)
};

struct SelectParams {
    int K, L, B;   // K chains per instance; the grid has one CTA per instance
    int M;         // instances (1 for the plain extend)
    int has_dest, compare_f32, k_offset;
    const double* dist;
    const int32_t* nok;
    const int32_t* goal_step;
    const float* states;
    const float* rest;
    const uint32_t* flags;
    const float* controls;
    mps_extend_best* best;
    float* best_states;
    float* best_controls;
    uint32_t* best_flags;
    // optional fused tree append
    int append, parent;
    long long* tree_size;      // [segments]
    size_t tree_cap;           // row stride (total columns)
    size_t tree_seg;           // columns per segment (= tree_cap for the plain tree)
    const int32_t* inst_ids;   // [M] segment of instance m, or nullptr (segment 0)
    const int32_t* parents;    // [M] parent node (local index) of instance m, or nullptr (= parent)
    float* tree_states;
    float* tree_controls;
    int32_t* tree_parent;
    int32_t* tree_slice;
    // M == 1 only: after everything else the CTA writes done_seq here (pinned host memory, system-scope fence first),
    // so that a caller whose outputs live in pinned memory can wait for the number instead of synchronising the stream
    volatile unsigned int* done_flag;
    unsigned int done_seq;
};

// exchange step of the multi-GPU extend: pick the global winner among the gathered per-rank records
struct MergeParams {
    const unsigned char* gathered;  // [world][rec_bytes] packed records in rank order
    int world, L, B;
    size_t rec_bytes;
    int has_dest, compare_f32;
    unsigned char* record;      // [rec_bytes] this ctx's record buffer: receives the winner's record
    mps_extend_best* head_out;  // optional second copy of the head (pinned host memory), or nullptr
    // optional tree append on the GPU that owns the tree
    int append, parent;
    long long* tree_size;
    size_t tree_cap;
    float* tree_states;
    float* tree_controls;
    int32_t* tree_parent;
    int32_t* tree_slice;
};

struct ValidityParams {
    const DevScene* scene;
    size_t scene_bytes, warp_bytes;
    const float* states;  // [n][3B]
    int n;
    uint8_t* valid;
    uint32_t* flags;
    // few states, one CTA, everything in pinned host memory: the CTA signs its results with done_seq (system-scope
    // fence) and the host waits for the number - no copy in either direction, no stream synchronise
    volatile unsigned int* done_flag;
    unsigned int done_seq;
};

size_t mps_rollout_warp_bytes(const DevScene& sc);
size_t mps_rollout_group_bytes(const DevScene& sc, int mcap);
size_t mps_rollout_scene_bytes();
// Tuning aids of the rollout launcher (0 = the launcher's own rule).  Read from the environment when a context is
// created, not per process: every context can be given its own, which is how the GPU tests force each build.
//   MPS_ROLLOUT_GROUP = 8 | 16 | 32 lanes per chain, MPS_ROLLOUT_MINB = 3 | 4 register budget of the unpacked build,
//   MPS_ROLLOUT_CTAS_PER_SM resident CTAs, MPS_ROLLOUT_ACTIVE_WARPS warps of a CTA that pull chains,
//   MPS_ROLLOUT_SOLO_BUILD = 0 no <4,1,32> build, MPS_ROLLED_MAX_BODIES largest scene that takes the rolled sweeps.
struct RolloutTuning {
    int group, minb, ctas_per_sm, active_warps, solo_build, rolled_max_bodies;
};
RolloutTuning mps_rollout_tuning_from_env();
cudaError_t mps_launch_rollout(const RolloutParams& p, const RolloutTuning& tune, int num_sms, int n_bodies,
                               cudaStream_t stream);
cudaError_t mps_launch_chain_order(const float* controls, const int32_t* n_ctrl, int K, int L, const float accel[3],
                                   float dt, int32_t* order, cudaStream_t stream);
cudaError_t mps_launch_select(const SelectParams& p, cudaStream_t stream);
cudaError_t mps_launch_merge_records(const MergeParams& p, cudaStream_t stream);
cudaError_t mps_launch_validity(const ValidityParams& p, int num_sms, cudaStream_t stream);
