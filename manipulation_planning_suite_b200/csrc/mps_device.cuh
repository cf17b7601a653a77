// mps_device.cuh — device-side scene layout and math helpers shared by the CUDA
// translation units of libmps_b200.so (sm_100a only).
//
// Arithmetic contract: the whole library is compiled with
// -fmad=false, so every +,-,*,/, sqrtf and every explicit fmaf below is one
// correctly rounded IEEE operation — the same operation sequence the physics
// spec mps-planar-v3 writes down (the spec says where a multiply-add is fused).  No libm sin/cos on the device path: mps_sincos is the spec's own
// polynomial.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mps_b200.h"

// dynamic shared memory of a kernel (the test-side SIMT emulator, tests/emu, substitutes a host buffer)
#ifdef MPS_CPU_EMU
#define MPS_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#else
#define MPS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define MPS_PI_D 3.14159265358979323846
#define MPS_PI_F 3.14159265358979323846f
#define MPS_TWO_PI_F (2.0f * 3.14159265358979323846f)
#define MPS_DBL_EPS 2.2204460492503131e-16

#define MPS_MAX_PAIRS (MPS_MAX_BODIES * (MPS_MAX_BODIES - 1) / 2 + MPS_MAX_BODIES * MPS_MAX_STATICS)

// Scene as the kernels see it: the mps_scene of the ABI plus the derived constants
// (inverse masses, friction caps, bounding radii, static rotations, pair table).
struct DevScene {
    int B, S, robot, vel_iters;
    int P;        // number of (body, body|static) pairs
    int Mcap;     // manifold capacity per rollout
    int strict;   // strict_active_contacts
    int slot_stride;  // bytes per row of the per-warp slot table (multiple of 4)
    int force_general;  // test hook (env MPS_FORCE_GENERAL_SWEEP=1 at context creation): always take the general sweep
    int jam_steps, _pad_i;
    float dt, t_max, slop, beta_dt, max_bias, pen_max, v_rest2, w_rest, ref_tol;
    float x_min, x_max, y_min, y_max;
    float accel[3];
    uint32_t static_forbidden;
    double pw, ow;
    int shape[MPS_MAX_BODIES];
    float hx[MPS_MAX_BODIES], hy[MPS_MAX_BODIES], brad[MPS_MAX_BODIES];
    float inv_m[MPS_MAX_BODIES], inv_i[MPS_MAX_BODIES], mass[MPS_MAX_BODIES], inertia[MPS_MAX_BODIES];
    float lin_cap[MPS_MAX_BODIES], ang_cap[MPS_MAX_BODIES], mu_c[MPS_MAX_BODIES];
    uint32_t pair_forbidden[MPS_MAX_BODIES];
    int s_shape[MPS_MAX_STATICS];
    float s_x[MPS_MAX_STATICS], s_y[MPS_MAX_STATICS], s_c[MPS_MAX_STATICS], s_s[MPS_MAX_STATICS];
    float s_hx[MPS_MAX_STATICS], s_hy[MPS_MAX_STATICS], s_brad[MPS_MAX_STATICS], s_mu[MPS_MAX_STATICS];
    uint16_t pairs[MPS_MAX_PAIRS];  // a | b << 8 ; b >= B is static b - B
};

// deterministic sin/cos of the spec: Cody-Waite reduction by pi/2 + fixed polynomials
__host__ __device__ __forceinline__ void mps_sincos(float a, float* s_out, float* c_out)
{
    const float two_over_pi = 0.63661977236758134f;
    const float p_hi = 1.5703125f;
    const float p_mid = 4.837512969970703125e-4f;
    const float p_lo = 7.54978995489188216e-8f;
    float kf = rintf(a * two_over_pi);
    int k = (int)kf;
    float r = fmaf(-kf, p_hi, a);
    r = fmaf(-kf, p_mid, r);
    r = fmaf(-kf, p_lo, r);
    float z = r * r;
    float sp = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    sp = sp * z;
    sp = fmaf(sp, r, r);
    float cp = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    cp = cp * z;
    cp = fmaf(cp, z, fmaf(-0.5f, z, 1.0f));
    int q = k & 3;
    float sn = (q == 0) ? sp : (q == 1) ? cp : (q == 2) ? -sp : -cp;
    float cs = (q == 0) ? cp : (q == 1) ? -sp : (q == 2) ? -cp : sp;
    *s_out = sn;
    *c_out = cs;
}

__host__ __device__ __forceinline__ float mps_wrap_angle(float th)
{
    if (th >= MPS_PI_F)
        th -= MPS_TWO_PI_F;
    else if (th < -MPS_PI_F)
        th += MPS_TWO_PI_F;
    return th;
}

// normalize_orientation(double&) of the reference (util/Math.cpp:8-16)
__host__ __device__ __forceinline__ double mps_normalize_orientation(double value)
{
    double v = fmod(value, 2.0 * MPS_PI_D);
    if (v < -MPS_PI_D)
        v += 2.0 * MPS_PI_D;
    else if (v >= MPS_PI_D)
        v -= 2.0 * MPS_PI_D;
    return v;
}

// per-body term of SimEnvWorldStateDistanceMeasure::distance in the reference's double arithmetic
__host__ __device__ __forceinline__ double mps_body_distance(float ax, float ay, float ath, float bx, float by,
                                                             float bth, double pw, double ow)
{
    double dx = (double)ax - (double)bx;
    double dy = (double)ay - (double)by;
    double dpos = sqrt(dx * dx + dy * dy);
    double d = fabs((double)ath - (double)bth);
    double arc = (d > MPS_PI_D) ? 2.0 * MPS_PI_D - d : d;
    return pw * dpos + ow * arc;
}

#define MPS_CUDA_CHECK(expr, ctx_)                                                        \
    do {                                                                                  \
        cudaError_t err__ = (expr);                                                       \
        if (err__ != cudaSuccess) return mps_fail_cuda((ctx_), err__, #expr, __LINE__);   \
    } while (0)
