// forest_step.h — sampling kernels of the device-resident NaiveRRT iteration (mps_forest_naive_step)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mps_b200.h"

struct NaiveSampleParams {
    int M, C, K, B;
    const int32_t* inst_ids;  // [M] local instance ids
    long long instance_offset;
    unsigned long long seed;
    int iteration;
    float x_min, x_max, y_min, y_max;
    float goal_bias, target_bias, robot_bias;
    int robot;
    const mps_goal* goals;  // device
    int n_goals;
    float v_limit, w_limit, dur_min, dur_max;
    float* cand;       // [M][C][3B]
    uint32_t* masks;   // [M] active-object mask of the distance
    float* controls;   // [M][K][5]
};

cudaError_t mps_launch_naive_sample(const NaiveSampleParams& p, cudaStream_t stream);
// samples[m] = first valid candidate of instance m (the last one when none is valid)
cudaError_t mps_launch_naive_pick(int M, int C, int B, const float* cand, const uint8_t* valid, float* samples,
                                  cudaStream_t stream);
// parents[m] = (int32) nearest[m]
cudaError_t mps_launch_idx_to_parent(int M, const long long* nearest, int32_t* parents, cudaStream_t stream);

// ---- action sequences of HybridActionRRT / OracleRRT generated on the device (mps_extend_upload_generated) --------
struct HybridGenParams {
    mps_hybrid_gen gen;
    int K, L, B, robot;
    float accel[3];
    const float* start;  // [3B] device
    const float* dest;   // [3B] device
    float* controls;     // [K][L][5] device
    int32_t* n_ctrl;     // [K]
    unsigned int* truncated;  // [1] chains that needed more than L controls
};
cudaError_t mps_launch_hybrid_controls(const HybridGenParams& p, cudaStream_t stream);
// n random controls (RampVelocityControlSampler::sample) of the device sampler, for tests of its distribution
cudaError_t mps_launch_sample_controls(unsigned long long seed, int iteration, long long gid, int n, float v_limit,
                                       float w_limit, float dur_min, float dur_max, float* controls, cudaStream_t stream);
