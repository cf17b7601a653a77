// forest_step.cu — the host side of a NaiveRRT iteration (RearrangementRRT::sample, sampleActiveObject,
// RampVelocityControlSampler::sample; pushing/algorithm/RRT.cpp:171-190,246-257, ompl/control/RampVelocityControl.cpp:353-394,
// ompl/state/goal/ObjectsRelocationGoal.cpp:67-102,133-136) for M independent planner instances at once, on the
// device, so that a lockstep multi-seed sweep needs no host random numbers and no uploads per iteration.
//
// Random numbers are counter-based: u(seed, iteration, global instance id, purpose, index) is a hash, so an
// instance's draws depend neither on the other instances nor on how the instances are split over GPUs.
#include "forest_step.h"

namespace {

enum { TAG_GOAL = 1, TAG_OBJ = 2, TAG_OBJ_UNIFORM = 3, TAG_TGT = 4, TAG_CAND = 5, TAG_BALL = 6, TAG_CTRL = 7 };

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// uniform in [0, 1) with 24 random bits
__device__ __forceinline__ float u01(unsigned long long seed, int iteration, long long gid, int tag, int idx)
{
    unsigned long long h = mix64(seed ^ mix64((unsigned long long)iteration * 0x100000001B3ull + (unsigned long long)tag));
    h = mix64(h ^ mix64((unsigned long long)gid));
    h = mix64(h ^ (unsigned long long)idx);
    return (float)(h >> 40) * (1.0f / 16777216.0f);
}

__global__ void naive_sample_kernel(NaiveSampleParams p)
{
    const int B = p.B;
    const float two_pi = 6.28318530717958647692f;
    // candidate states: thread = (instance, candidate)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.M * p.C; t += gridDim.x * blockDim.x) {
        const int m = t / p.C, c = t - m * p.C;
        const long long gid = p.instance_offset + p.inst_ids[m];
        const bool is_goal = u01(p.seed, p.iteration, gid, TAG_GOAL, 0) < p.goal_bias;
        float* out = p.cand + (size_t)t * 3 * B;
        for (int b = 0; b < B; ++b) {
            const int i = (c * B + b) * 3;
            out[3 * b] = p.x_min + (p.x_max - p.x_min) * u01(p.seed, p.iteration, gid, TAG_CAND, i);
            out[3 * b + 1] = p.y_min + (p.y_max - p.y_min) * u01(p.seed, p.iteration, gid, TAG_CAND, i + 1);
            out[3 * b + 2] = -3.14159265358979323846f + two_pi * u01(p.seed, p.iteration, gid, TAG_CAND, i + 2);
        }
        if (is_goal) {  // ObjectsRelocationGoal::sampleGoal: every target inside its goal disc
            for (int g = 0; g < p.n_goals; ++g) {
                const int b = p.goals[g].body;
                const float r = p.goals[g].tol * sqrtf(u01(p.seed, p.iteration, gid, TAG_BALL, (c * p.n_goals + g) * 2));
                float sn, cs;
                __sincosf(two_pi * u01(p.seed, p.iteration, gid, TAG_BALL, (c * p.n_goals + g) * 2 + 1), &sn, &cs);
                out[3 * b] = p.goals[g].x + r * cs;
                out[3 * b + 1] = p.goals[g].y + r * sn;
            }
        }
        if (c == 0) {
            // active object: the goal's target when the sample is a goal, else sampleActiveObject (RRT.cpp:246-257);
            // the last target is twice as likely (ObjectsRelocationGoal.cpp:133-136)
            int pick = (int)(u01(p.seed, p.iteration, gid, TAG_TGT, 0) * (float)(p.n_goals + 1));
            if (pick > p.n_goals - 1) pick = p.n_goals - 1;
            const int target = p.n_goals > 0 ? p.goals[pick].body : 0;
            const float uo = u01(p.seed, p.iteration, gid, TAG_OBJ, 0);
            int obj = (int)(u01(p.seed, p.iteration, gid, TAG_OBJ_UNIFORM, 0) * (float)B);
            if (obj > B - 1) obj = B - 1;
            if (uo < p.target_bias)
                obj = target;
            else if (uo < p.target_bias + p.robot_bias)
                obj = p.robot;
            if (is_goal) obj = target;
            p.masks[m] = 1u << obj;
        }
    }
    // controls: thread = (instance, control sample); (vx, vy) uniform in the disc of radius v_limit
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.M * p.K; t += gridDim.x * blockDim.x) {
        const int m = t / p.K, k = t - m * p.K;
        const long long gid = p.instance_offset + p.inst_ids[m];
        const float r = p.v_limit * sqrtf(u01(p.seed, p.iteration, gid, TAG_CTRL, k * 4));
        float sn, cs;
        __sincosf(two_pi * u01(p.seed, p.iteration, gid, TAG_CTRL, k * 4 + 1), &sn, &cs);
        float* c = p.controls + (size_t)t * MPS_CONTROL_DIM;
        c[0] = r * cs;
        c[1] = r * sn;
        c[2] = (2.0f * u01(p.seed, p.iteration, gid, TAG_CTRL, k * 4 + 2) - 1.0f) * p.w_limit;
        c[3] = p.dur_min + (p.dur_max - p.dur_min) * u01(p.seed, p.iteration, gid, TAG_CTRL, k * 4 + 3);
        c[4] = 0.0f;
    }
}

__global__ void naive_pick_kernel(int M, int C, int dim, const float* __restrict__ cand,
                                  const uint8_t* __restrict__ valid, float* __restrict__ samples)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < M * dim; t += gridDim.x * blockDim.x) {
        const int m = t / dim, d = t - m * dim;
        int first = C - 1;
        for (int c = 0; c < C; ++c)
            if (valid[m * C + c]) {
                first = c;
                break;
            }
        samples[t] = cand[((size_t)m * C + first) * dim + d];
    }
}

__global__ void idx_to_parent_kernel(int M, const long long* __restrict__ nearest, int32_t* __restrict__ parents)
{
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x)
        parents[m] = nearest[m] < 0 ? 0 : (int32_t)nearest[m];
}

int grid_of(long long n)
{
    long long g = (n + 255) / 256;
    if (g > 1184) g = 1184;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

cudaError_t mps_launch_naive_sample(const NaiveSampleParams& p, cudaStream_t stream)
{
    const long long n = (long long)p.M * (p.C > p.K ? p.C : p.K);
    naive_sample_kernel<<<grid_of(n), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t mps_launch_naive_pick(int M, int C, int B, const float* cand, const uint8_t* valid, float* samples,
                                  cudaStream_t stream)
{
    naive_pick_kernel<<<grid_of((long long)M * 3 * B), 256, 0, stream>>>(M, C, 3 * B, cand, valid, samples);
    return cudaGetLastError();
}

cudaError_t mps_launch_idx_to_parent(int M, const long long* nearest, int32_t* parents, cudaStream_t stream)
{
    idx_to_parent_kernel<<<grid_of(M), 256, 0, stream>>>(M, nearest, parents);
    return cudaGetLastError();
}
