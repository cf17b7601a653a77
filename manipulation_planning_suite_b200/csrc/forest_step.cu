// forest_step.cu — the host side of a NaiveRRT iteration (RearrangementRRT::sample, sampleActiveObject,
// RampVelocityControlSampler::sample; pushing/algorithm/RRT.cpp:171-190,246-257, ompl/control/RampVelocityControl.cpp:353-394,
// ompl/state/goal/ObjectsRelocationGoal.cpp:67-102,133-136) for M independent planner instances at once, on the
// device, so that a lockstep multi-seed sweep needs no host random numbers and no uploads per iteration.
//
// Random numbers are counter-based: u(seed, iteration, global instance id, purpose, index) is a hash, so an
// instance's draws depend neither on the other instances nor on how the instances are split over GPUs.
#include "forest_step.h"
#include "mps_device.cuh"

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

// RampVelocityControlSampler::sample (ompl/control/RampVelocityControl.cpp:353-394): (vx, vy) uniform in the disc of
// radius v_limit (the reference normalises a gaussian pair and scales by v_limit * u^(1/2); radius v_limit * sqrt(u)
// at a uniform angle is the same distribution), omega and the plateau duration uniform, rest time 0
__device__ __forceinline__ void sample_control(unsigned long long seed, int iteration, long long gid, int tag, int idx,
                                               float v_limit, float w_limit, float dur_min, float dur_max, float* c)
{
    const float two_pi = 6.28318530717958647692f;
    const float r = v_limit * sqrtf(u01(seed, iteration, gid, tag, idx * 4));
    float sn, cs;
    __sincosf(two_pi * u01(seed, iteration, gid, tag, idx * 4 + 1), &sn, &cs);
    c[0] = r * cs;
    c[1] = r * sn;
    c[2] = (2.0f * u01(seed, iteration, gid, tag, idx * 4 + 2) - 1.0f) * w_limit;
    c[3] = dur_min + (dur_max - dur_min) * u01(seed, iteration, gid, tag, idx * 4 + 3);
    c[4] = 0.0f;
}

__global__ void sample_controls_kernel(unsigned long long seed, int iteration, long long gid, int n, float v_limit,
                                       float w_limit, float dur_min, float dur_max, float* controls)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        sample_control(seed, iteration, gid, TAG_CTRL, t, v_limit, w_limit, dur_min, dur_max, controls + (size_t)t * MPS_CONTROL_DIM);
}

// ---- oracle action sequences --------------------------------------------------------------------------------------
enum { TAG_H_DIE = 11, TAG_H_OBJ = 12, TAG_H_OBJ_UNIFORM = 13, TAG_H_TGT = 14, TAG_H_GAUSS = 15, TAG_H_ZERO = 16, TAG_H_RANDOM = 17 };

__device__ __forceinline__ float shortest_so2(float a, float b)  // util/Math.cpp:33-44
{
    float v = b - a;
    if (fabsf(v) > 3.14159265358979323846f) v = v > 0.0f ? v - 6.28318530717958647692f : v + 6.28318530717958647692f;
    return v;
}

// RampVelocityControl::computeRamp's acceleration time for plateau velocities v (RampVelocityControl.cpp:98-118)
__device__ __forceinline__ float accel_time_of(float vx, float vy, float vw, const float acc[3])
{
    float t = fabsf(vx / acc[0]);
    t = fmaxf(t, fabsf(vy / acc[1]));
    t = fmaxf(t, fabsf(vw / acc[2]));
    return t;
}

__device__ __forceinline__ float norm3(float x, float y, float z, float s)
{
    return sqrtf(s * x * s * x + s * y * s * y + s * z * s * z);
}

// RampComputer::steer (pushing/oracle/RampComputer.cpp:40-92): ramp controls that take the robot from (cx, cy, cth) to
// (dx_, dy_, dth) along the straight line of its configuration space; emit(n, vx, vy, w, duration) receives the n-th
// control; returns how many controls the steer consists of
template <class Emit>
__device__ __forceinline__ int steer(float cx, float cy, float cth, float dx_, float dy_, float dth, const float vlim[3],
                                     const float acc[3], float dur_max, Emit emit)
{
    float ex = dx_ - cx, ey = dy_ - cy, ez = shortest_so2(cth, dth);
    float distance = sqrtf(ex * ex + ey * ey + ez * ez);
    if (!(distance > 0.0f)) return 0;
    ex /= distance;
    ey /= distance;
    ez /= distance;
    float vabs = 3.402823466e+38f;
    vabs = fminf(vabs, vlim[0] / fabsf(ex));
    vabs = fminf(vabs, vlim[1] / fabsf(ey));
    vabs = fminf(vabs, vlim[2] / fabsf(ez));
    float vx = vabs * ex, vy = vabs * ey, vw = vabs * ez;
    float ta = accel_time_of(vx, vy, vw, acc);
    float dist_accel = norm3(vx, vy, vw, ta);
    int n = 0;
    while (distance > 0.0f && n < 64) {
        if (dist_accel > distance) {  // the acceleration phases alone would overshoot: a slower plateau velocity
            vabs = distance / ta;
            vx = vabs * ex;
            vy = vabs * ey;
            vw = vabs * ez;
            ta = accel_time_of(vx, vy, vw, acc);
            dist_accel = norm3(vx, vy, vw, ta);
        }
        distance = distance - dist_accel;
        const float max_plateau = fminf(norm3(vx, vy, vw, dur_max), distance);
        emit(n, vx, vy, vw, max_plateau / vabs);
        ++n;
        distance = distance - max_plateau;
    }
    return n;
}

// standard normal from two uniforms (Box-Muller)
__device__ __forceinline__ float gauss01(float u1, float u2)
{
    float sn, cs;
    mps_sincos(6.28318530717958647692f * u2, &sn, &cs);  // the spec's own sincos: no libm slow path (and its stack frame)
    return sqrtf(-2.0f * logf(1.0f - u1)) * cs;
}

// One thread per chain: HybridActionRRT::sampleActionSequence (pushing/algorithm/RRT.cpp:491-529) with
// OracleControlSampler::{steerRobot, steerPush, sampleFeasibleState, randomControl} (oracle/OracleControlSampler.cpp:
// 77-207) and HumanOracle::{sampleFeasibleState, predictAction} (oracle/HumanOracle.cpp:111-159) behind them.
__global__ void hybrid_controls_kernel(HybridGenParams p)
{
    const mps_hybrid_gen& g = p.gen;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < p.K; k += gridDim.x * blockDim.x) {
        float* out = p.controls + (size_t)k * p.L * MPS_CONTROL_DIM;
        for (int i = 0; i < p.L * MPS_CONTROL_DIM; ++i) out[i] = 0.0f;
        const long long gid = (long long)g.k_offset + k;
        int n = 0;
        const int L = p.L;
        auto write_control = [out, L](int i, float vx, float vy, float vw, float duration) {
            if (i < L) {
                float* c = out + i * MPS_CONTROL_DIM;
                c[0] = vx;
                c[1] = vy;
                c[2] = vw;
                c[3] = duration;
                c[4] = 0.0f;
            }
        };
        const float die = u01(g.seed, g.iteration, gid, TAG_H_DIE, 0);
        const float cur_r[3] = {p.start[3 * p.robot], p.start[3 * p.robot + 1], p.start[3 * p.robot + 2]};
        if (die < g.action_randomness) {
            sample_control(g.seed, g.iteration, gid, TAG_H_RANDOM, 0, g.velocity_limits[0], g.velocity_limits[2], g.duration_min,
                           g.duration_max, out);
            n = 1;
        } else {
            // which object to move: the caller's (OracleRRT), else sampleActiveObject (RRT.cpp:246-257)
            int obj = g.fixed_object;
            if (obj < 0) {
                const float uo = u01(g.seed, g.iteration, gid, TAG_H_OBJ, 0);
                obj = (int)(u01(g.seed, g.iteration, gid, TAG_H_OBJ_UNIFORM, 0) * (float)p.B);
                if (obj > p.B - 1) obj = p.B - 1;
                if (uo < g.target_bias && g.n_targets > 0) {
                    int pick = (int)(u01(g.seed, g.iteration, gid, TAG_H_TGT, 0) * (float)g.n_targets);
                    if (pick > g.n_targets - 1) pick = g.n_targets - 1;
                    obj = g.targets[pick];
                } else if (uo < g.target_bias + g.robot_bias) {
                    obj = p.robot;
                }
            }
            if (obj == p.robot) {  // transit primitive
                n = steer(cur_r[0], cur_r[1], cur_r[2], p.dest[3 * p.robot], p.dest[3 * p.robot + 1], p.dest[3 * p.robot + 2],
                          g.velocity_limits, p.accel, g.duration_max, write_control);
            } else {  // transfer primitive: to a feasible pushing state, then the push (the world is assumed unchanged)
                const float cur_o[3] = {p.start[3 * obj], p.start[3 * obj + 1], p.start[3 * obj + 2]};
                const float des_o[3] = {p.dest[3 * obj], p.dest[3 * obj + 1], p.dest[3 * obj + 2]};
                // HumanOracle::sampleFeasibleState
                const float pushing_dist = g.optimal_push_distance + g.push_distance_tolerance *
                    gauss01(u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 0), u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 1));
                const float angle = g.push_angle_tolerance *
                    gauss01(u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 2), u01(g.seed, g.iteration, gid, TAG_H_GAUSS, 3));
                float rel[3] = {des_o[0] - cur_o[0], des_o[1] - cur_o[1], shortest_so2(cur_o[2], des_o[2])};
                if (sqrtf(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]) == 0.0f) {
                    rel[0] = 0.0001f * (0.5f - u01(g.seed, g.iteration, gid, TAG_H_ZERO, 0));
                    rel[1] = 0.0001f * (0.5f - u01(g.seed, g.iteration, gid, TAG_H_ZERO, 1));
                }
                const float nrm = sqrtf(rel[0] * rel[0] + rel[1] * rel[1]);
                float pd[2] = {1.0f, 0.0f};
                if (nrm > 0.0f) {
                    pd[0] = rel[0] / nrm;
                    pd[1] = rel[1] / nrm;
                }
                float ca, sa;
                mps_sincos(angle, &sa, &ca);
                float feas[3];
                feas[0] = cur_o[0] - pushing_dist * (ca * pd[0] - sa * pd[1]);
                feas[1] = cur_o[1] - pushing_dist * (sa * pd[0] + ca * pd[1]);
                feas[2] = atan2f(pd[1], pd[0]) - 1.57f;  // the reference's floating end-effector (HumanOracle.cpp:156)
                // setConfiguration normalises the SO(2) component (SimEnvState.cpp:131-135)
                double th = (double)feas[2];  // in [-pi - 1.57, pi - 1.57]: fmod(th, 2 pi) of the normalisation is th itself
                if (th < -MPS_PI_D)
                    th += 2.0 * MPS_PI_D;
                else if (th >= MPS_PI_D)
                    th -= 2.0 * MPS_PI_D;
                feas[2] = (float)th;
                n = steer(cur_r[0], cur_r[1], cur_r[2], feas[0], feas[1], feas[2], g.velocity_limits, p.accel, g.duration_max,
                          write_control);
                // HumanOracle::predictAction from the feasible state: the first control of the steer towards the
                // object's destination, half an object size short of it; a null control when the object does not translate
                float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
                const float r2x = des_o[0] - cur_o[0], r2y = des_o[1] - cur_o[1];
                const float n2 = sqrtf(r2x * r2x + r2y * r2y);
                if (n2 != 0.0f) {
                    const float half = g.object_size[obj] / 2.0f;
                    steer(feas[0], feas[1], feas[2], des_o[0] - r2x / n2 * half, des_o[1] - r2y / n2 * half, feas[2],
                          g.velocity_limits, p.accel, g.duration_max, [&](int i, float a, float b, float c, float d) {
                              if (i == 0) {
                                  p0 = a;
                                  p1 = b;
                                  p2 = c;
                                  p3 = d;
                              }
                          });
                }
                write_control(n, p0, p1, p2, p3);
                ++n;
            }
        }
        if (n > p.L) {
            atomicAdd(p.truncated, 1u);
            n = p.L;
        }
        p.n_ctrl[k] = n;
    }
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
        sample_control(p.seed, p.iteration, gid, TAG_CTRL, k, p.v_limit, p.w_limit, p.dur_min, p.dur_max,
                       p.controls + (size_t)t * MPS_CONTROL_DIM);
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

cudaError_t mps_launch_hybrid_controls(const HybridGenParams& p, cudaStream_t stream)
{
    hybrid_controls_kernel<<<grid_of(p.K), 128, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t mps_launch_sample_controls(unsigned long long seed, int iteration, long long gid, int n, float v_limit,
                                       float w_limit, float dur_min, float dur_max, float* controls, cudaStream_t stream)
{
    sample_controls_kernel<<<grid_of(n), 256, 0, stream>>>(seed, iteration, gid, n, v_limit, w_limit, dur_min, dur_max, controls);
    return cudaGetLastError();
}
