// rollout.cu — K push rollouts of the planar robot/object contact simulation with the
// validity / goal / slice-ball tests fused in, one lane group per rollout chain (sm_100a).
//
// What it replaces in the reference (paths relative to the reference root):
//   the K-sample loop of NaiveControlSampler::sampleTo      ompl/control/NaiveControlSampler.cpp:46-64
//   the K-chain loop of HybridActionRRT::extend            pushing/algorithm/RRT.cpp:458-477,531-565
//   SimEnvStatePropagator::propagate                       ompl/control/SimEnvStatePropagator.cpp:42-118
//   RampVelocityControl::{computeRamp,getVelocity}          ompl/control/RampVelocityControl.cpp:98-158
//   SimEnvValidityChecker::isValid + CollisionPolicy       ompl/state/SimEnvState.cpp:1365-1394,1441-1502
//   SimEnvWorldStateDistanceMeasure::distance              ompl/state/SimEnvWorldStateDistanceMeasure.cpp:34-51
//   ObjectsRelocationGoal::distanceGoal / isSatisfied      ompl/state/goal/ObjectsRelocationGoal.cpp:108-123
// The physics step itself follows the spec mps-planar-v3.
//
// Mapping to the machine:
//   * one lane group (a whole warp, or 8 / 16 lanes for small scenes in launches that fill the machine) = one
//     rollout chain; lane i of the group owns body i: pose and twist live in that lane's registers, poses are
//     mirrored to shared memory for the pair tests;
//   * broad phase: lane = pair index, ballot-compacted into a candidate list that carries a skin and is reused
//     until a body has moved too far; narrow phase: lane = listed pair (exact reach test, then SAT / clipping),
//     touching pairs compacted into manifold records of 8 float4 in the group's shared slice;
//   * solver: Gauss-Seidel sweeps in the spec's canonical order (colour r < B holds the body pairs with
//     (i + j) mod B == r, colour B + s the pairs (i, static s)); manifolds that share no dynamic body commute, so
//     every manifold is put in the earliest round after the manifolds it depends on and a round's manifolds are
//     visited at the same time by their owner lanes, the partner's twist arriving by shuffle.  A sequential CPU
//     sweep in canonical order gives the same bits;
//   * rest test / policy test: group votes;
//   * groups pull chain indices from an atomic counter (rollout lengths vary by 10x).
#include "mps_device.cuh"
#include "rollout.h"

#include <atomic>
#include <cstdlib>

#define FULL 0xffffffffu

// A rollout chain is run by a group of G consecutive lanes of a warp (G = 32: the whole warp; G = 16 / 8: two /
// four chains per warp for scenes with at most G bodies, so that the lanes a small scene leaves idle carry
// other chains).  `lane` in the device functions below is the lane WITHIN the group (= body / pair index);
// every warp collective uses the group's mask, so groups of one warp may be in different places of the code.
struct Ln {
    int g;       // lane within the group
    int base;    // first lane of the group in the warp
    unsigned m;  // member mask of the group
    __device__ __forceinline__ int abs_lane(int body) const { return base + body; }
    __device__ __forceinline__ unsigned below() const { return (1u << (base + g)) - 1u; }  // lanes before this one
};
template <int G>
__device__ __forceinline__ Ln make_ln()
{
    Ln ln;
    const int pl = threadIdx.x & 31;
    ln.g = pl & (G - 1);
    ln.base = pl & ~(G - 1);
    ln.m = G == 32 ? FULL : (((1u << G) - 1u) << ln.base);
    return ln;
}

namespace {


// One manifold = 8 float4 (128 B) in the group's shared slice, read and written with 128-bit accesses:
//   [0] meta (a | b << 8 | count << 16), nx, ny, friction coefficient sqrt(mu_a mu_b)
//   [1] point 0: rna, rnb, rta, rtb        (lever arms r x n, r x t of the two bodies)
//   [2] point 0: normal mass, tangent mass, normal mass * bias, coupling t0-t1
//   [3] point 1: rna, rnb, rta, rtb
//   [4] point 1: normal mass, tangent mass, normal mass * bias, coupling n0-n1
//   [5] couplings t0-n0, t0-n1, t1-n0, t1-n1
//   [6] accumulated impulses (n0, t0, n1, t1) of owner lane a   [7] the same, owner lane b's private copy
// A missing second point carries zero rows.
#ifndef MF_VEC4
#define MF_VEC4 8
#endif
// sweeps of at most this many rounds keep the manifold rows in registers (see physics_step): two in the build
// with the large register budget, MPS_REG_ROUNDS in the packed 128-register builds, none in the unpacked
// 128-register build (measured on B200: +6 % on cfg2, +8 % packed at B = 7, -5 % unpacked at B = 17)
#ifndef MPS_REG_ROUNDS
#define MPS_REG_ROUNDS 1
#endif
// ... and in the build for launches of one CTA per SM (MINB = 1: up to 255 registers)
#ifndef MPS_REG_ROUNDS_SOLO
#define MPS_REG_ROUNDS_SOLO 4
#endif
#define MF_FIELDS (4 * MF_VEC4)
// Unrolling of the iteration loop of the register-resident sweeps.  The iterations are serial (Gauss-Seidel), so
// unrolling only removes the loop's branch and costs code the step has to fetch: measured on B200
// (profiles/r02_variants_shortcuts_coldcalls.txt) against the compiler's own choice (about 2): not unrolled, cfg 2
// +0.7 %, two chains per warp +6 .. 9 %, four chains per warp -7 % (kept at 2 there); unrolled 4 / 8 times cfg 2 loses
// 7 / 25 %.
template <int G>
struct SweepUnroll {
    static constexpr int value = G == 8 ? 2 : 1;
};

struct Shape {
    int shape;
    float x, y, c, s, hx, hy;
};

struct Manifold {
    int count;
    float nx, ny;
    float rax[2], ray[2], rbx[2], rby[2], sep[2];
};

__device__ __forceinline__ int collide_disc_disc(const Shape& A, const Shape& Bs, Manifold& m, float& gap)
{
    float dx = Bs.x - A.x;
    float dy = Bs.y - A.y;
    float d2 = fmaf(dx, dx, dy * dy);
    float R = A.hx + Bs.hx;
    if (d2 > R * R) {
        gap = sqrtf(d2) - R;
        return 0;
    }
    float d = sqrtf(d2);
    float nx, ny;
    if (d > 1e-9f) {
        nx = dx / d;
        ny = dy / d;
    } else {
        nx = 1.0f;
        ny = 0.0f;
    }
    float sep = d - R;
    float h = fmaf(0.5f, sep, A.hx);
    m.nx = nx;
    m.ny = ny;
    m.count = 1;
    m.rax[0] = nx * h;
    m.ray[0] = ny * h;
    m.rbx[0] = m.rax[0] - dx;
    m.rby[0] = m.ray[0] - dy;
    m.sep[0] = sep;
    return 1;
}

// box X against disc D; a_is_box: body a of the pair is the box
__device__ __forceinline__ int collide_box_disc(const Shape& X, const Shape& D, bool a_is_box, Manifold& m,
                                                float& gap)
{
    float ddx = D.x - X.x;
    float ddy = D.y - X.y;
    float lx = fmaf(ddx, X.c, ddy * X.s);
    float ly = fmaf(ddy, X.c, -(ddx * X.s));
    float qx = fminf(fmaxf(lx, -X.hx), X.hx);
    float qy = fminf(fmaxf(ly, -X.hy), X.hy);
    float r = D.hx;
    float nlx, nly, sep, spx, spy;
    if (qx == lx && qy == ly) {
        float px = X.hx - fabsf(lx);
        float py = X.hy - fabsf(ly);
        if (px < py) {
            nlx = lx >= 0.0f ? 1.0f : -1.0f;
            nly = 0.0f;
            sep = -(px + r);
            spx = nlx * X.hx;
            spy = ly;
        } else {
            nlx = 0.0f;
            nly = ly >= 0.0f ? 1.0f : -1.0f;
            sep = -(py + r);
            spx = lx;
            spy = nly * X.hy;
        }
    } else {
        float vx = lx - qx;
        float vy = ly - qy;
        float d2 = fmaf(vx, vx, vy * vy);
        if (d2 > r * r) {
            gap = sqrtf(d2) - r;
            return 0;
        }
        float d = sqrtf(d2);
        nlx = vx / d;
        nly = vy / d;
        sep = d - r;
        spx = qx;
        spy = qy;
    }
    float hs = 0.5f * sep;
    float cplx = fmaf(nlx, hs, spx);
    float cply = fmaf(nly, hs, spy);
    float rxx = fmaf(X.c, cplx, -(X.s * cply));
    float rxy = fmaf(X.s, cplx, X.c * cply);
    float nx = fmaf(X.c, nlx, -(X.s * nly));
    float ny = fmaf(X.s, nlx, X.c * nly);
    float rdx = rxx - ddx;
    float rdy = rxy - ddy;
    m.count = 1;
    m.sep[0] = sep;
    if (a_is_box) {
        m.nx = nx;
        m.ny = ny;
        m.rax[0] = rxx;
        m.ray[0] = rxy;
        m.rbx[0] = rdx;
        m.rby[0] = rdy;
    } else {
        m.nx = -nx;
        m.ny = -ny;
        m.rax[0] = rdx;
        m.ray[0] = rdy;
        m.rbx[0] = rxx;
        m.rby[0] = rxy;
    }
    return 1;
}

// rectangle-rectangle: 4-axis SAT, reference face = least penetration (A unless B is better by
// ref_tol), incident edge clipped to the reference face's side planes
__device__ __forceinline__ int collide_box_box(const Shape& A, const Shape& Bs, float ref_tol, Manifold& m,
                                               float& gap)
{
    float dx = Bs.x - A.x;
    float dy = Bs.y - A.y;
    float d00 = fmaf(A.c, Bs.c, A.s * Bs.s);
    float d01 = fmaf(A.s, Bs.c, -(A.c * Bs.s));
    float ad00 = fabsf(d00);
    float ad01 = fabsf(d01);
    float pax = fmaf(dx, A.c, dy * A.s);
    float pay = fmaf(dy, A.c, -(dx * A.s));
    float pbx = fmaf(dx, Bs.c, dy * Bs.s);
    float pby = fmaf(dy, Bs.c, -(dx * Bs.s));
    float sep_ax = fabsf(pax) - A.hx - fmaf(Bs.hx, ad00, Bs.hy * ad01);
    float sep_ay = fabsf(pay) - A.hy - fmaf(Bs.hx, ad01, Bs.hy * ad00);
    float sep_bx = fabsf(pbx) - Bs.hx - fmaf(A.hx, ad00, A.hy * ad01);
    float sep_by = fabsf(pby) - Bs.hy - fmaf(A.hx, ad01, A.hy * ad00);
    if (sep_ax > 0.0f || sep_ay > 0.0f || sep_bx > 0.0f || sep_by > 0.0f) {
        gap = fmaxf(fmaxf(sep_ax, sep_ay), fmaxf(sep_bx, sep_by));  // separated by at least this along an axis
        return 0;
    }
    bool a_face_y = sep_ay > sep_ax;
    bool b_face_y = sep_by > sep_bx;
    float sep_a = a_face_y ? sep_ay : sep_ax;
    float sep_b = b_face_y ? sep_by : sep_bx;
    bool ref_is_b = sep_b > sep_a + ref_tol;
    float Rc = ref_is_b ? Bs.c : A.c, Rs = ref_is_b ? Bs.s : A.s;
    float Rhx = ref_is_b ? Bs.hx : A.hx, Rhy = ref_is_b ? Bs.hy : A.hy;
    float Ic = ref_is_b ? A.c : Bs.c, Is = ref_is_b ? A.s : Bs.s;
    float Ihx = ref_is_b ? A.hx : Bs.hx, Ihy = ref_is_b ? A.hy : Bs.hy;
    float ddx = ref_is_b ? -dx : dx;
    float ddy = ref_is_b ? -dy : dy;
    bool face_y = ref_is_b ? b_face_y : a_face_y;
    float ux = face_y ? -Rs : Rc;
    float uy = face_y ? Rc : Rs;
    float hu = face_y ? Rhy : Rhx;
    float ht = face_y ? Rhx : Rhy;
    float proj = fmaf(ddx, ux, ddy * uy);
    float sgn = proj >= 0.0f ? 1.0f : -1.0f;
    float nrx = sgn * ux;
    float nry = sgn * uy;
    float trx = -nry;
    float tryy = nrx;
    float e0 = fmaf(nrx, Ic, nry * Is);
    float e1 = fmaf(nry, Ic, -(nrx * Is));
    bool inc_y = !(fabsf(e0) >= fabsf(e1));
    float sm, mx, my, hm, he;
    if (!inc_y) {
        sm = e0 > 0.0f ? -1.0f : 1.0f;
        mx = sm * Ic;
        my = sm * Is;
        hm = Ihx;
        he = Ihy;
    } else {
        sm = e1 > 0.0f ? -1.0f : 1.0f;
        mx = sm * -Is;
        my = sm * Ic;
        hm = Ihy;
        he = Ihx;
    }
    float cx = fmaf(mx, hm, ddx);
    float cy = fmaf(my, hm, ddy);
    float ex = -my;
    float ey = mx;
    float v1x = fmaf(ex, he, cx);
    float v1y = fmaf(ey, he, cy);
    float v2x = fmaf(-ex, he, cx);
    float v2y = fmaf(-ey, he, cy);
    float t1 = fmaf(v1x, trx, v1y * tryy);
    float t2 = fmaf(v2x, trx, v2y * tryy);
    float s1 = fmaf(v1x, nrx, v1y * nry) - hu;
    float s2 = fmaf(v2x, nrx, v2y * nry) - hu;
    if ((t1 > ht && t2 > ht) || (t1 < -ht && t2 < -ht)) return 0;
    // clip the incident edge to the side planes, without branches: the clipped values are computed for
    // both ends (a zero denominator only occurs when nothing is clipped) and selected
    const bool hi1 = t1 > ht, lo1 = !hi1 && t1 < -ht;
    const bool hi2 = t2 > ht, lo2 = !hi2 && t2 < -ht;
    const float lim1 = hi1 ? ht : -ht;
    const float lim2 = hi2 ? ht : -ht;
    const float al1 = (lim1 - t1) / (t2 - t1);
    const float al2 = (lim2 - t2) / (t1 - t2);
    const float cl1 = fmaf(al1, s2 - s1, s1);
    const float cl2 = fmaf(al2, s1 - s2, s2);
    const bool c1 = hi1 || lo1, c2 = hi2 || lo2;
    float ct0 = c1 ? lim1 : t1, cs0 = c1 ? cl1 : s1;
    float ct1 = c2 ? lim2 : t2, cs1 = c2 ? cl2 : s2;
    // contact points: the clipped ends that lie behind the reference face, midway between the two surfaces;
    // both are computed, the second moves to slot 0 when the first is not a contact
    float qax[2], qay[2], qbx[2], qby[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float csk = k == 0 ? cs0 : cs1;
        const float ctk = k == 0 ? ct0 : ct1;
        const float h = fmaf(0.5f, csk, hu);
        const float rrx = fmaf(trx, ctk, nrx * h);
        const float rry = fmaf(tryy, ctk, nry * h);
        const float rix = rrx - ddx;
        const float riy = rry - ddy;
        qax[k] = ref_is_b ? rix : rrx;
        qay[k] = ref_is_b ? riy : rry;
        qbx[k] = ref_is_b ? rrx : rix;
        qby[k] = ref_is_b ? rry : riy;
    }
    const bool v0 = cs0 <= 0.0f, v1 = cs1 <= 0.0f;
    const int n = (v0 ? 1 : 0) + (v1 ? 1 : 0);
    m.rax[0] = v0 ? qax[0] : qax[1];
    m.ray[0] = v0 ? qay[0] : qay[1];
    m.rbx[0] = v0 ? qbx[0] : qbx[1];
    m.rby[0] = v0 ? qby[0] : qby[1];
    m.sep[0] = v0 ? cs0 : cs1;
    m.rax[1] = qax[1];
    m.ray[1] = qay[1];
    m.rbx[1] = qbx[1];
    m.rby[1] = qby[1];
    m.sep[1] = cs1;
    if (n == 0) return 0;
    m.count = n;
    m.nx = ref_is_b ? -nrx : nrx;
    m.ny = ref_is_b ? -nry : nry;
    return n;
}

// gap: when the shapes do not touch, a lower bound of their distance (0 when none is known)
__device__ __forceinline__ int collide(const Shape& A, const Shape& Bs, float ref_tol, Manifold& m, float& gap)
{
    gap = 0.0f;
    if (A.shape == MPS_SHAPE_BOX && Bs.shape == MPS_SHAPE_BOX) return collide_box_box(A, Bs, ref_tol, m, gap);
    if (A.shape == MPS_SHAPE_BOX && Bs.shape == MPS_SHAPE_DISC) return collide_box_disc(A, Bs, true, m, gap);
    if (A.shape == MPS_SHAPE_DISC && Bs.shape == MPS_SHAPE_BOX) return collide_box_disc(Bs, A, false, m, gap);
    return collide_disc_disc(A, Bs, m, gap);
}

__device__ __forceinline__ unsigned char* opaque_shared_base(unsigned char* dyn)
{
#ifdef MPS_CPU_EMU
    return dyn;
#else
    unsigned int s32 = (unsigned int)__cvta_generic_to_shared(dyn);
    asm volatile("" : "+r"(s32));
    return (unsigned char*)__cvta_shared_to_generic((size_t)s32);
#endif
}

// per-chain working set in shared memory
struct WarpMem {
    float4* pose;     // [B + S] (x, y, cos, sin): the bodies' poses mirrored for the pair tests, then the statics'
    float* mv;        // [B + S] upper bound of the distance any point of the body has travelled in this chain
                      // (statics: 0)
    float4* man;      // [Mcap][MF_VEC4]
    uint8_t* slot;    // [B][slot_stride]: manifold slot of pair (row body, column body|static); never
                      // cleared — an entry is trusted only if the manifold it names carries the same pair
    uint16_t* cand;   // [P] broad-phase survivors of the current step
    uint16_t* sig;    // [Mcap] pair of every manifold slot the round schedule was built for (packed groups)
    int mcap;         // manifold slots of this group (sc.Mcap, or the small capacity of a packed launch)
};

// poses / travelled distances of bodies and statics share one index space (a static is B + k)
__host__ __device__ __forceinline__ int mps_pose_slots(int B, int S) { return (B + S + 3) & ~3; }

__device__ __forceinline__ WarpMem carve_warp_mem(const DevScene& sc, unsigned char* wbase, int Mcap)
{
    WarpMem W;
    const int np = mps_pose_slots(sc.B, sc.S);
    W.pose = reinterpret_cast<float4*>(wbase);
    W.mv = reinterpret_cast<float*>(W.pose + np);
    W.man = reinterpret_cast<float4*>(W.mv + np);
    W.slot = reinterpret_cast<uint8_t*>(W.man + MF_VEC4 * Mcap);
    W.cand = reinterpret_cast<uint16_t*>(W.slot + ((sc.B * sc.slot_stride + 3) & ~3));
    W.sig = W.cand + ((sc.P + 1) & ~1);
    W.mcap = Mcap;
    return W;
}

// the statics' entries of the pose table (once per chain group; the bodies' entries follow every step)
__device__ __forceinline__ void publish_statics(const DevScene& sc, const WarpMem& W, const Ln ln, int G)
{
    for (int k = ln.g; k < sc.S; k += G) {
        W.pose[sc.B + k] = make_float4(sc.s_x[k], sc.s_y[k], sc.s_c[k], sc.s_s[k]);
        W.mv[sc.B + k] = 0.0f;
    }
    __syncwarp(ln.m);
}

// State a chain carries from one physics step to the next (registers).
//  * Candidate list with a skin: the broad phase keeps every pair within reach + MPS_SKIN and is rebuilt only
//    when a body has moved more than MPS_SKIN_MOVE from where it was at the last rebuild (bounding circles do
//    not care about rotation).  Each step re-applies the exact reach test to the listed pairs, so the pairs that
//    enter the narrow phase are the same as with a fresh broad phase.
//  * Round schedule of the solver, reused while the list of touching pairs is unchanged.
#define MPS_SKIN 0.04f
#define MPS_SKIN_MOVE2 (0.018f * 0.018f) /* (0.45 * MPS_SKIN)^2: two bodies together stay below the skin */
struct StepCache {
    float rx, ry;  // this lane's body position at the last rebuild
    int n_list;    // pairs in W.cand, < 0: none
    float clear_until;  // list entry `lane` was found separated by a gap: it cannot touch while the distance
                        // its two bodies have travelled (W.mv) stays below this; -inf = unknown
    unsigned ent[4];
    int rounds, nm;       // nm < 0: no schedule
    unsigned sig0, sig1;  // pair of manifold slot `lane` / `lane + 32` the schedule was built for
    __device__ __forceinline__ void reset()
    {
        n_list = -1;
        nm = -1;
        clear_until = -1.0f;
    }
};

struct DetectResult {
    bool bad;       // a touching pair the policy forbids
    bool overflow;  // more touching pairs than Mcap
    float max_pen;
    int nm;         // manifolds stored (<= Mcap)
    unsigned candidates, touching;
};

__device__ __forceinline__ void load_pair_shapes(const DevScene& sc, const WarpMem& W, int a, int b, Shape& A,
                                                 Shape& Bs)
{
    const int B = sc.B;
    const float4 pa = W.pose[a], pb = W.pose[b];
    A.shape = sc.shape[a];
    A.x = pa.x;
    A.y = pa.y;
    A.c = pa.z;
    A.s = pa.w;
    A.hx = sc.hx[a];
    A.hy = sc.hy[a];
    // the second shape is a body or a static (pose table index B + k): the addresses are selected, not the
    // code path
    const bool body = b < B;
    const int k = body ? b : b - B;
    Bs.shape = *(body ? &sc.shape[k] : &sc.s_shape[k]);
    Bs.x = pb.x;
    Bs.y = pb.y;
    Bs.c = pb.z;
    Bs.s = pb.w;
    Bs.hx = *(body ? &sc.hx[k] : &sc.s_hx[k]);
    Bs.hy = *(body ? &sc.hy[k] : &sc.s_hy[k]);
}

// reach test of one pair (bounding circles; body circle against the static's own box / disc - walls are long,
// their bounding circle would make every body a candidate).  skin = 0 is the exact test of the narrow phase's
// entry; conservative with a small margin for statics; only a filter.
__device__ __forceinline__ bool pair_in_reach(const DevScene& sc, const WarpMem& W, unsigned pr, float skin)
{
    const int B = sc.B;
    const int a = pr & 0xff, b = pr >> 8;
    const float4 pa = W.pose[a], pb = W.pose[b];
    if (b < B) {
        float rr = sc.brad[a] + sc.brad[b];
        rr = rr + skin;
        float dx = pb.x - pa.x;
        float dy = pb.y - pa.y;
        return !(dx * dx + dy * dy > rr * rr);
    }
    const int k = b - B;
    float dx = pa.x - pb.x;
    float dy = pa.y - pb.y;
    float rr = sc.brad[a] * 1.001f + 1e-5f;
    if (sc.s_shape[k] == MPS_SHAPE_BOX) {
        rr = rr + skin;
        float lx = dx * pb.z + dy * pb.w;
        float ly = dy * pb.z - dx * pb.w;
        float ex = fmaxf(fabsf(lx) - sc.s_hx[k], 0.0f);
        float ey = fmaxf(fabsf(ly) - sc.s_hy[k], 0.0f);
        return !(ex * ex + ey * ey > rr * rr);
    }
    rr += sc.s_hx[k];
    rr = rr + skin;
    return !(dx * dx + dy * dy > rr * rr);
}

// broad phase: lane = pair of the scene's table, bounding circles with a skin, survivors compacted into W.cand
// in table order (= the solver's canonical order)
template <int G>
__device__ __forceinline__ void rebuild_list(const DevScene& sc, const WarpMem& W, const Ln ln, StepCache& cache)
{
    const int lane = ln.g;
    // two chunks of G pairs per trip so that their loads and tests overlap
    int nl = 0;
    const unsigned lt = ln.below();
    for (int base = 0; base < sc.P; base += 2 * G) {
        const int p0 = base + lane, p1 = base + G + lane;
        unsigned pr0 = 0, pr1 = 0;
        bool c0 = false, c1 = false;
        if (p0 < sc.P) {
            pr0 = sc.pairs[p0];
            c0 = pair_in_reach(sc, W, pr0, MPS_SKIN);
        }
        if (p1 < sc.P) {
            pr1 = sc.pairs[p1];
            c1 = pair_in_reach(sc, W, pr1, MPS_SKIN);
        }
        const unsigned m0 = __ballot_sync(ln.m, c0);
        const unsigned m1 = __ballot_sync(ln.m, c1);
        const int n0 = __popc(m0);
        if (c0) W.cand[nl + __popc(m0 & lt)] = (uint16_t)pr0;
        if (c1) W.cand[nl + n0 + __popc(m1 & lt)] = (uint16_t)pr1;
        nl += n0 + __popc(m1);
    }
    cache.n_list = nl;
    cache.clear_until = -1.0f;
    if (lane < sc.B) {
        const float4 me = W.pose[lane];
        cache.rx = me.x;
        cache.ry = me.y;
    }
    __syncwarp(ln.m);
}

// manifold record of a touching pair in the chain's shared slice.  prepare: with the solver rows (layout: 0
// meta|n|friction, 1 / 3 lever arms of point 0 / 1, 2 / 4 inverse effective masses, bias term and the t-t / n-n
// coupling, 5 t-n couplings, 6 / 7 accumulated impulses, one private copy per owner lane); the rows of a
// missing second point are all zero and run through the same arithmetic.  any_two: some manifold of this trip
// has a second point (otherwise its arithmetic is left out: the rows stay zero).
__device__ __forceinline__ void store_manifold(const DevScene& sc, float4* mf, const Manifold& m, int cnt, int a, int b,
                                               bool any_two, bool prepare)
{
    const int B = sc.B;
    const bool bstat = b >= B;
    const int meta = a | (b << 8) | (cnt << 16);
    if (!prepare) {
        mf[0] = make_float4(__int_as_float(meta), m.nx, m.ny, 0.0f);
        return;
    }
    const float ma = sc.inv_m[a], ia = sc.inv_i[a];
    const int kb_ = bstat ? b - B : b;
    const float mbq = sc.inv_m[bstat ? 0 : b], ibq = sc.inv_i[bstat ? 0 : b];
    const float mb = bstat ? 0.0f : mbq, ib = bstat ? 0.0f : ibq;
    const float mub = *(bstat ? &sc.s_mu[kb_] : &sc.mu_c[kb_]);
    float tx = m.ny, ty = -m.nx;
    float iarna[2] = {0.0f, 0.0f}, ibrnb[2] = {0.0f, 0.0f}, iarta[2] = {0.0f, 0.0f}, ibrtb[2] = {0.0f, 0.0f};
    float rna[2] = {0.0f, 0.0f}, rnb[2] = {0.0f, 0.0f}, rta[2] = {0.0f, 0.0f}, rtb[2] = {0.0f, 0.0f};
    float nmass[2] = {0.0f, 0.0f}, tmass[2] = {0.0f, 0.0f}, kb[2] = {0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        // both points go through the arithmetic; a missing second point is zeroed afterwards
        if (k == 1 && !any_two) break;
        const bool has = k < cnt;
        const float rax = m.rax[k], ray = m.ray[k], rbx = m.rbx[k], rby = m.rby[k];
        const float sepk = m.sep[k];
        rna[k] = has ? fmaf(rax, m.ny, -(ray * m.nx)) : 0.0f;
        rnb[k] = has ? fmaf(rbx, m.ny, -(rby * m.nx)) : 0.0f;
        rta[k] = has ? fmaf(rax, ty, -(ray * tx)) : 0.0f;
        rtb[k] = has ? fmaf(rbx, ty, -(rby * tx)) : 0.0f;
        iarna[k] = ia * rna[k];
        ibrnb[k] = ib * rnb[k];
        iarta[k] = ia * rta[k];
        ibrtb[k] = ib * rtb[k];
        const float kn = fmaf(ibrnb[k], rnb[k], fmaf(iarna[k], rna[k], ma + mb));
        const float kt = fmaf(ibrtb[k], rtb[k], fmaf(iarta[k], rta[k], ma + mb));
        const float pen_k = -sepk - sc.slop;
        const float bias = pen_k > 0.0f ? fminf(sc.beta_dt * pen_k, sc.max_bias) : 0.0f;
        const float rn = 1.0f / kn, rt = 1.0f / kt;
        nmass[k] = (has && kn > 0.0f) ? rn : 0.0f;
        tmass[k] = (has && kt > 0.0f) ? rt : 0.0f;
        kb[k] = nmass[k] * bias;
    }
    const bool two = cnt > 1;
    const float ktn00 = fmaf(iarta[0], rna[0], ibrtb[0] * rnb[0]);
    const float ktt_ = fmaf(ibrtb[0], rtb[1], fmaf(iarta[0], rta[1], ma + mb));
    const float knn_ = fmaf(ibrnb[0], rnb[1], fmaf(iarna[0], rna[1], ma + mb));
    const float ktn01_ = fmaf(iarta[0], rna[1], ibrtb[0] * rnb[1]);
    const float ktn10_ = fmaf(iarta[1], rna[0], ibrtb[1] * rnb[0]);
    const float ktn11_ = fmaf(iarta[1], rna[1], ibrtb[1] * rnb[1]);
    const float ktt = two ? ktt_ : 0.0f, knn = two ? knn_ : 0.0f;
    const float ktn01 = two ? ktn01_ : 0.0f, ktn10 = two ? ktn10_ : 0.0f, ktn11 = two ? ktn11_ : 0.0f;
    mf[1] = make_float4(rna[0], rnb[0], rta[0], rtb[0]);
    mf[2] = make_float4(nmass[0], tmass[0], kb[0], ktt);
    mf[3] = make_float4(rna[1], rnb[1], rta[1], rtb[1]);
    mf[4] = make_float4(nmass[1], tmass[1], kb[1], knn);
    mf[5] = make_float4(ktn00, ktn01, ktn10, ktn11);
    mf[6] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    mf[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    mf[0] = make_float4(__int_as_float(meta), m.nx, m.ny, sqrtf(sc.mu_c[a] * mub));
}

// collision detection at the poses in W.pose.  prepare: also build the solver rows.
template <int G>
__device__ __forceinline__ DetectResult detect(const DevScene& sc, const WarpMem& W, const Ln ln, bool prepare,
                                               StepCache& cache, long long* t_mid = nullptr,
                                               long long* sub = nullptr)
{
    const int lane = ln.g;
    const int B = sc.B;
    const int Mcap = W.mcap;
    // ---- broad phase: lane = pair, bounding circles, survivors compacted into W.cand ------------
    // reach test of one pair; skin = 0 is the exact test of the physics spec
    auto pair_near = [&](unsigned pr, float skin) -> bool { return pair_in_reach(sc, W, pr, skin); };
    bool rebuild = cache.n_list < 0;
    if (!rebuild) {
        bool moved = false;
        if (lane < B) {
            const float4 me = W.pose[lane];
            float dx = me.x - cache.rx;
            float dy = me.y - cache.ry;
            moved = dx * dx + dy * dy > MPS_SKIN_MOVE2;
        }
        rebuild = __any_sync(ln.m, moved);
    }
    if (rebuild) rebuild_list<G>(sc, W, ln, cache);
    const int nlist = cache.n_list;
    int ncand = 0;
#ifdef MPS_PHASE_PROF
    if (t_mid) *t_mid = clock64();
#endif
    // ---- narrow phase: lane = listed pair; pairs out of exact reach drop out first ----------------
    int nm = 0;
    bool lane_bad = false, lane_over = false;
    float lane_pen = 0.0f;
    unsigned n_touch = 0;
    for (int base = 0; base < nlist; base += G) {
#ifdef MPS_PHASE_PROF
        long long tsb = clock64();
#endif
        int c = base + lane;
        Manifold m = {};
        int cnt = 0, a = 0, b = 0;
        // A pair found separated by a gap g cannot touch before its bodies have travelled g between them:
        // until then it is not looked at again (entries of the first trip only).  Exact: the skipped test
        // could only have said "not touching".  `candidates` counts the pairs that are looked at.
        bool look = false;
        float travelled = 0.0f;
        if (c < nlist) {
            unsigned pr = W.cand[c];
            a = pr & 0xff;
            b = pr >> 8;
            travelled = W.mv[a] + W.mv[b];  // a static's entry is 0
            look = base != 0 || !(travelled < cache.clear_until);
        }
        if (!__any_sync(ln.m, look)) continue;
        const bool near = look && pair_near((unsigned)a | ((unsigned)b << 8), 0.0f);
        ncand += __popc(__ballot_sync(ln.m, near));
#ifdef MPS_PHASE_PROF
        long long ts0 = clock64();
        if (sub) sub[0] += ts0 - tsb;
#endif
        if (near) {
            Shape A, Bs;
            float gap;
            load_pair_shapes(sc, W, a, b, A, Bs);
            cnt = collide(A, Bs, sc.ref_tol, m, gap);
            if (base == 0) cache.clear_until = gap > 4e-5f ? travelled + (0.9f * gap - 2e-5f) : -1.0f;
        }
        bool touch = cnt > 0;
        unsigned tmask = __ballot_sync(ln.m, touch);
        // no manifold of this trip has a second point (most trips): its rows stay zero without the arithmetic
        const bool any_two = G < 32 || __any_sync(ln.m, cnt > 1);
#ifdef MPS_PHASE_PROF
        long long ts1 = clock64();
        if (sub) sub[1] += ts1 - ts0;
#endif
        if (tmask == 0u) continue;
        n_touch += __popc(tmask);
#ifdef MPS_PHASE_PROF
        struct RowTimer {
            long long t0, *dst;
            __device__ ~RowTimer() { if (dst) dst[2] += clock64() - t0; }
        } row_timer{ts1, sub};
#endif
        if (touch) {
            float pen = -m.sep[0];
            if (pen > lane_pen) lane_pen = pen;
            if (cnt > 1) {
                pen = -m.sep[1];
                if (pen > lane_pen) lane_pen = pen;
            }
            const bool bstat = b >= B;
            const unsigned forbid = bstat ? (sc.static_forbidden >> a) : (sc.pair_forbidden[a] >> (b & 31));
            if (forbid & 1u) lane_bad = true;
            int idx = nm + __popc(tmask & ln.below());
            if (idx < Mcap) {
                store_manifold(sc, W.man + idx * MF_VEC4, m, cnt, a, b, any_two, prepare);
            } else {
                lane_over = true;
            }
        }
        nm += __popc(tmask);
    }
    DetectResult r;
    // nothing touches (most steps): no lane has a flag, a penetration or a manifold to publish.  Only in the packed
    // builds (four chains per warp, 7 bodies, K = 65 536: 19.4 -> 20.7 M rollouts/s); one chain per warp loses 1 .. 3 %
    // with the extra branch (profiles/r02_variants_shortcuts_coldcalls.txt)
    if (G < 32 && n_touch == 0u) {
        r.bad = false;
        r.overflow = false;
        r.max_pen = 0.0f;
        r.nm = 0;
        r.candidates = (unsigned)ncand;
        r.touching = 0u;
        return r;
    }
    const unsigned fl = __reduce_or_sync(ln.m, (lane_bad ? 1u : 0u) | (lane_over ? 2u : 0u));
    r.bad = (fl & 1u) != 0u;
    r.overflow = (fl & 2u) != 0u;
    r.max_pen = __uint_as_float(__reduce_max_sync(ln.m, __float_as_uint(lane_pen)));
    r.nm = nm < Mcap ? nm : Mcap;
    r.candidates = (unsigned)ncand;
    r.touching = n_touch;
    __syncwarp(ln.m);
    return r;
}

// manifold slot of body `lane` in colour `col`, or 0xff.  The slot table is never cleared: an entry counts
// only if it names a live manifold that carries exactly this pair.
__device__ __forceinline__ int lookup_slot(const DevScene& sc, const WarpMem& W, int lane, int col, int nm)
{
    const int B = sc.B;
    if (lane >= B) return 0xff;
    int a = lane, b = col;
    if (col < B) {
        int j = col - lane;
        if (j < 0) j += B;
        if (j == lane) return 0xff;
        a = lane < j ? lane : j;
        b = lane < j ? j : lane;
        col = j;
    }
    int sl = W.slot[lane * sc.slot_stride + col];
    if (sl >= nm) return 0xff;
    int meta = __float_as_int(W.man[sl * MF_VEC4].x);
    return (meta & 0xffff) == (a | (b << 8)) ? sl : 0xff;
}

// registers of the lane that owns body `lane`
struct Body {
    float x, y, th, c, s, vx, vy, w;
    float inv_m, inv_i, mass, inertia, lin_cap, ang_cap;
    float mv, brad;  // travelled-distance bound (see WarpMem::mv), bounding radius
    bool dynamic;  // lane < B and not the robot
    bool live;     // lane < B
};

__device__ __forceinline__ void publish_pose(const WarpMem& W, const Body& b, const Ln ln)
{
    const int lane = ln.g;
    if (b.live) {
        W.pose[lane] = make_float4(b.x, b.y, b.c, b.s);
        W.mv[lane] = b.mv;
    }
    __syncwarp(ln.m);
}

__device__ __forceinline__ void set_state(const WarpMem& W, Body& b, const Ln ln, float x, float y, float th)
{
    if (b.live) {
        b.x = x;
        b.y = y;
        b.th = mps_wrap_angle(th);
        mps_sincos(b.th, &b.s, &b.c);
        b.mv += 1e-5f;  // the pose handed in may differ from the integrated one by a rounding of theta
    }
    b.vx = 0.0f;
    b.vy = 0.0f;
    b.w = 0.0f;
    publish_pose(W, b, ln);
}

struct StepStats {
    unsigned steps, candidates, touching, colour_passes;
#ifdef MPS_PHASE_PROF
    long long ph[11];  // broad phase, narrow phase, round scheduling, solver rounds, friction + integration,
                       // contact steps, narrow: reach test / collide / rows
#endif
};
#ifdef MPS_PHASE_PROF
#define PH_T(var) long long var = clock64()
#define PH_ADD(i, a, b) st.ph[i] += (b) - (a)
#else
#define PH_T(var)
#define PH_ADD(i, a, b)
#endif

// One pass of the Gauss-Seidel sweep: every lane that owns a body with a manifold scheduled in this pass
// solves it (both owners of a body pair do, redundantly and identically; the partner's twist arrives by
// shuffle) and updates its own body.  Per lane: sl = manifold slot or 0xff, j = partner lane (self when none
// or static), stat = the partner is a static.
// The pass has no data-dependent branch and no barrier: lanes without a manifold run the arithmetic on slot 0
// and drop the result, a one-point manifold carries zero rows for the second point, and each owner lane keeps
// its own copy of the accumulated impulses, so no lane reads what another lane writes during the sweep.
// rows of one manifold as the sweep reads them (WarpMem::man layout)
struct MRows {
    float4 m0, r0, g0, r1, g1, kc;
};

__device__ __forceinline__ MRows load_rows(const float4* mf)
{
    MRows R;
    R.m0 = mf[0];
    R.r0 = mf[1];
    R.g0 = mf[2];
    R.r1 = mf[3];
    R.g1 = mf[4];
    R.kc = mf[5];
    return R;
}

// the arithmetic of one visit; acc = this owner lane's copy of the accumulated impulses.
// TWO = false: every manifold visited in this pass has a single point.  The rows of the missing second point are
// all zero, so its impulses are zeros and every term they enter adds a zero: leaving that arithmetic out gives
// the same values (only the sign of a result that is exactly zero may differ, which no later operation can
// turn into a different value: the step uses +, -, *, fma, min, max and comparisons only).
template <bool TWO>
__device__ __forceinline__ void solve_core(const MRows& R, float4& acc, Body& bd, const Ln ln, bool act, int j,
                                           bool stat, bool is_a)
{
    float ovx = __shfl_sync(ln.m, bd.vx, ln.abs_lane(j));
    float ovy = __shfl_sync(ln.m, bd.vy, ln.abs_lane(j));
    float ow_ = __shfl_sync(ln.m, bd.w, ln.abs_lane(j));
    if (stat) {
        ovx = 0.0f;
        ovy = 0.0f;
        ow_ = 0.0f;
    }
    const float4 m0 = R.m0, r0 = R.r0, g0 = R.g0, r1 = R.r1, g1 = R.g1, kc = R.kc;
    const float nx = m0.y, ny = m0.z, fr = m0.w;
    const float tx = ny, ty = -nx;
    // lever arms of this lane's own body
    const float on0 = is_a ? r0.x : r0.y, ot0 = is_a ? r0.z : r0.w;
    const float on1 = is_a ? r1.x : r1.y, ot1 = is_a ? r1.z : r1.w;
    const float sm = is_a ? -bd.inv_m : bd.inv_m;
    const float si = is_a ? -bd.inv_i : bd.inv_i;
    const float vax = is_a ? bd.vx : ovx, vay = is_a ? bd.vy : ovy, wa = is_a ? bd.w : ow_;
    const float vbx = is_a ? ovx : bd.vx, vby = is_a ? ovy : bd.vy, wb = is_a ? ow_ : bd.w;
    // relative velocity along every row, from the twists at the start of the visit
    const float dvx = vbx - vax;
    const float dvy = vby - vay;
    const float ut0 = fmaf(dvx, tx, fmaf(dvy, ty, fmaf(wb, r0.w, -(wa * r0.z))));
    const float un0 = fmaf(dvx, nx, fmaf(dvy, ny, fmaf(wb, r0.y, -(wa * r0.x))));
    // friction first (Box2D order), point 0 then point 1, then non-penetration; an impulse on an earlier
    // row reaches a later row through the coupling coefficient
    float dt0, dt1 = 0.0f, dn0, dn1 = 0.0f;
    {
        const float lim = fr * acc.x;
        const float dl = g0.y * -ut0;
        dt0 = fminf(fmaxf(dl, -lim - acc.y), lim - acc.y);
        acc.y = acc.y + dt0;
    }
    if (TWO) {
        const float ut1 = fmaf(dvx, tx, fmaf(dvy, ty, fmaf(wb, r1.w, -(wa * r1.z))));
        const float u = fmaf(g0.w, dt0, ut1);
        const float lim = fr * acc.z;
        const float dl = g1.y * -u;
        dt1 = fminf(fmaxf(dl, -lim - acc.w), lim - acc.w);
        acc.w = acc.w + dt1;
    }
    {
        const float u = TWO ? fmaf(kc.z, dt1, fmaf(kc.x, dt0, un0)) : fmaf(kc.x, dt0, un0);
        const float dl = fmaf(-g0.x, u, g0.z);
        dn0 = fmaxf(dl, -acc.x);
        acc.x = acc.x + dn0;
    }
    if (TWO) {
        const float un1 = fmaf(dvx, nx, fmaf(dvy, ny, fmaf(wb, r1.y, -(wa * r1.x))));
        const float u = fmaf(g1.w, dn0, fmaf(kc.w, dt1, fmaf(kc.y, dt0, un1)));
        const float dl = fmaf(-g1.x, u, g1.z);
        dn1 = fmaxf(dl, -acc.z);
        acc.z = acc.z + dn1;
    }
    // total impulse on b (a receives the opposite); each lane updates its own body only
    const float st = TWO ? dt0 + dt1 : dt0;
    const float rr0 = fmaf(on0, dn0, ot0 * dt0);
    const float rr = TWO ? fmaf(on1, dn1, fmaf(ot1, dt1, rr0)) : rr0;
    const float px0 = fmaf(nx, dn0, tx * st);
    const float py0 = fmaf(ny, dn0, ty * st);
    const float px = TWO ? fmaf(nx, dn1, px0) : px0;
    const float py = TWO ? fmaf(ny, dn1, py0) : py0;
    const float nvx = fmaf(sm, px, bd.vx);
    const float nvy = fmaf(sm, py, bd.vy);
    const float nw = fmaf(si, rr, bd.w);
    if (act) {
        bd.vx = nvx;
        bd.vy = nvy;
        bd.w = nw;
    }
}

template <bool TWO = true>
__device__ __forceinline__ void solve_pass(const WarpMem& W, Body& bd, const Ln ln, int sl, int j, bool stat)
{
    const int lane = ln.g;
    const bool act = sl != 0xff;
    const bool is_a = stat || lane < j;
    float4* mf = W.man + (act ? sl : 0) * MF_VEC4;
    const MRows R = load_rows(mf);
    float4* accp = mf + (is_a ? 6 : 7);
    float4 acc = *accp;
    solve_core<TWO>(R, acc, bd, ln, act, j, stat, is_a);
    if (act) *accp = acc;
}

// Round schedule of the solver for the manifold list in W.man[0 .. nm) (canonical order): every manifold goes
// to the earliest round after the earlier manifolds that share a dynamic body with it,
//     round(m) = 1 + max(last round of body a, last round of body b);
// each lane packs its <= 4 entries as round << 16 | static << 15 | partner << 8 | slot (a body with more sends
// the step to the general sweep: rounds = -1).  With at most 4 rounds the entries are indexed by round.
__device__ __forceinline__ void build_schedule(const DevScene& sc, const WarpMem& W, const Ln ln, int nm,
                                               StepCache& cache)
{
    const int lane = ln.g;
    const int B = sc.B;
    unsigned e0 = 0u, e1 = 0u, e2 = 0u, e3 = 0u;
    int last = 0, ne = 0;
    // the manifold list is in canonical order (the pair table is): walk it once
    for (int m = 0; m < nm; ++m) {
        const int meta = __float_as_int(W.man[m * MF_VEC4].x);
        const int a = meta & 0xff, b = (meta >> 8) & 0xff;
        const bool stc = b >= B;
        // the robot is kinematic: its twist is only read, so manifolds that share nothing but the
        // robot commute and the robot neither waits nor is waited for
        const bool dyn_a = a != sc.robot;
        const bool dyn_b = !stc && b != sc.robot;
        int r = 0;
        if (dyn_a) r = __shfl_sync(ln.m, last, ln.abs_lane(a));
        if (dyn_b) {
            const int lb = __shfl_sync(ln.m, last, ln.abs_lane(b));
            r = r > lb ? r : lb;
        }
        r += 1;
        if ((dyn_a && lane == a) || (dyn_b && lane == b)) {
            last = r;
            const unsigned e = ((unsigned)r << 16) | (stc ? 0x8000u : 0u) |
                               ((unsigned)(stc ? lane : (lane == a ? b : a)) << 8) | (unsigned)m;
            e3 = ne == 3 ? e : e3;
            e2 = ne == 2 ? e : e2;
            e1 = ne == 1 ? e : e1;
            e0 = ne == 0 ? e : e0;
            ++ne;
        }
    }
    int nr = __reduce_max_sync(ln.m, last);
    if (__any_sync(ln.m, ne > 4) || sc.force_general) nr = -1;
    if (nr >= 0 && nr <= 4) {
        // few rounds (the usual case): index the entries by round, ent[r - 1] = this lane's
        // manifold of round r or 0, so the sweep needs no search
        unsigned b0 = 0u, b1 = 0u, b2 = 0u, b3 = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned e = c == 0 ? e0 : c == 1 ? e1 : c == 2 ? e2 : e3;
            const int er = (int)(e >> 16);
            b0 = er == 1 ? e : b0;
            b1 = er == 2 ? e : b1;
            b2 = er == 3 ? e : b2;
            b3 = er == 4 ? e : b3;
        }
        e0 = b0;
        e1 = b1;
        e2 = b2;
        e3 = b3;
    }
    cache.ent[0] = e0;
    cache.ent[1] = e1;
    cache.ent[2] = e2;
    cache.ent[3] = e3;
    cache.rounds = nr;
}

// General sweep, colour by colour in canonical order, for steps in which some body has more than 4
// manifolds (rare): the slot table and the set of active colours are built from the manifold list first.
template <int G>
__device__ __forceinline__ void general_sweep(const DevScene& sc, const WarpMem& W, Body& bd, const Ln ln, int nm,
                                              StepStats& st)
{
    const int lane = ln.g;
    const int B = sc.B;
    unsigned long long lane_col = 0ull;
    for (int m = lane; m < nm; m += G) {
        const int meta = __float_as_int(W.man[m * MF_VEC4].x);
        const int a = meta & 0xff, b = (meta >> 8) & 0xff;
        W.slot[a * sc.slot_stride + b] = (uint8_t)m;
        int col;
        if (b < B) {
            W.slot[b * sc.slot_stride + a] = (uint8_t)m;
            col = a + b;
            if (col >= B) col -= B;
        } else {
            col = b;
        }
        lane_col |= 1ull << col;
    }
    const unsigned lo = __reduce_or_sync(ln.m, (unsigned)(lane_col & 0xffffffffull));
    const unsigned hi = __reduce_or_sync(ln.m, (unsigned)(lane_col >> 32));
    const unsigned long long colours = ((unsigned long long)hi << 32) | lo;
    __syncwarp(ln.m);
    st.colour_passes += (unsigned)(__popcll(colours) * sc.vel_iters);
    for (int it = 0; it < sc.vel_iters; ++it) {
        unsigned long long cm = colours;
        while (cm) {
            const int col = __ffsll((long long)cm) - 1;
            cm &= cm - 1ull;
            const int slot = lookup_slot(sc, W, lane, col, nm);
            const bool stat = col >= B;
            int j = lane;
            if (!stat && bd.live) {
                j = col - lane;
                if (j < 0) j += B;
            }
            solve_pass(W, bd, ln, slot, j, stat);
        }
    }
}

// One round's manifold of this lane as the register-resident sweeps hold it (see physics_step)
struct RoundRegs {
    MRows R;
    float4 acc;
    int j;
    bool act, st, isa, two;
};

__device__ __forceinline__ RoundRegs load_round(const WarpMem& W, unsigned e, int lane)
{
    RoundRegs r;
    r.act = e != 0u;
    r.st = (e & 0x8000u) != 0u;
    r.j = r.act ? (int)((e >> 8) & 0x3fu) : lane;
    r.isa = r.st || lane < r.j;
    r.R = load_rows(W.man + (r.act ? (int)(e & 0xffu) : 0) * MF_VEC4);
    r.acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    r.two = r.act && ((__float_as_int(r.R.m0.x) >> 16) & 0xff) > 1;
    return r;
}

// vel_iters sweeps over NR rounds, rows and accumulated impulses in registers
template <bool TWO, int NR>
__device__ __forceinline__ void sweep_in_registers(RoundRegs (&r)[NR], Body& bd, const Ln ln, int iters)
{
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NR; ++k) solve_core<TWO>(r[k].R, r[k].acc, bd, ln, r[k].act, r[k].j, r[k].st, r[k].isa);
    }
}

// one physics step.  bad / overflow / max_pen come back warp-uniform.
template <int G, int RR, bool ROLLED>
__device__ __forceinline__ void physics_step(const DevScene& sc, const WarpMem& W, Body& bd, const Ln ln, float ux,
                                             float uy, float uw, bool& bad, bool& overflow, float& max_pen,
                                             StepStats& st, StepCache& cache)
{
    constexpr int kSweepUnroll = SweepUnroll<G>::value;
    const int lane = ln.g;
    const int B = sc.B;
    if (lane == sc.robot) {
        bd.vx = ux;
        bd.vy = uy;
        bd.w = uw;
    }
    PH_T(t0);
#ifdef MPS_PHASE_PROF
    long long t_mid = t0;
    DetectResult dr = detect<G>(sc, W, ln, true, cache, &t_mid, &st.ph[6]);
#else
    DetectResult dr = detect<G>(sc, W, ln, true, cache);
#endif
    PH_T(t1);
    PH_ADD(0, t0, t_mid);
    PH_ADD(1, t_mid, t1);
    bad = dr.bad;
    overflow = dr.overflow;
    max_pen = dr.max_pen;
    st.steps += 1;
    st.candidates += dr.candidates;
    st.touching += dr.touching;
    if (overflow) return;
    // support-plane friction, once per step before the contact iterations
    if (bd.dynamic) {
        // an impulse within the cap brings the body to rest exactly (spec v3 §3.2 step 3)
        const float imp = -bd.inertia * bd.w;
        const float impc = fminf(fmaxf(imp, -bd.ang_cap), bd.ang_cap);
        bd.w = impc == imp ? 0.0f : fmaf(bd.inv_i, impc, bd.w);
        const float ix = -bd.mass * bd.vx;
        const float iy = -bd.mass * bd.vy;
        const float n2 = fmaf(ix, ix, iy * iy);
        const float cap = bd.lin_cap;
        if (n2 > cap * cap) {
            const float sc_ = cap / sqrtf(n2);
            bd.vx = fmaf(bd.inv_m, ix * sc_, bd.vx);
            bd.vy = fmaf(bd.inv_m, iy * sc_, bd.vy);
        } else {
            bd.vx = 0.0f;
            bd.vy = 0.0f;
        }
    }
    if (dr.nm > 0) {
        {
            // Common case: no body has more than 4 manifolds.  The canonical order is colour by colour, but two
            // manifolds that share no body commute, so each manifold is scheduled in the earliest ROUND
            // after every earlier manifold (in canonical order) that shares a body with it:
            //     round(m) = 1 + max(last round of body a, last round of body b).
            // Any schedule that keeps the relative order of manifolds sharing a body produces the bits of
            // the canonical sweep; independent contacts (different islands) now solve concurrently.
            // Each lane packs its <= 4 entries as round << 16 | static << 15 | partner << 8 | slot; a body
            // with more manifolds than that sends the step to the general sweep below (rounds = -1).
            // The schedule depends only on which pair sits in which manifold slot: reuse the last one while
            // that list is unchanged (contact sets persist over many steps).
            unsigned sig0 = 0xffffffffu, sig1 = 0xffffffffu;
            bool same;
            bool two_step = true;  // some manifold of the step has a second point (unpacked 128-register build)
            if (G == 32) {
                if (lane < dr.nm) sig0 = (unsigned)(__float_as_int(W.man[lane * MF_VEC4].x) & 0xffff);
                if (lane + 32 < dr.nm) sig1 = (unsigned)(__float_as_int(W.man[(lane + 32) * MF_VEC4].x) & 0xffff);
                same = __all_sync(ln.m, cache.nm == dr.nm && sig0 == cache.sig0 && sig1 == cache.sig1);
                if (RR < 2) {
                    // does any manifold of the step have a second point?  (the build that keeps two rounds in
                    // registers asks the rows it has loaded instead; slots beyond 64 are not looked at: such a
                    // step takes the two-point code)
                    int cnt2 = 0;
                    if (lane < dr.nm) cnt2 = (__float_as_int(W.man[lane * MF_VEC4].x) >> 16) & 0xff;
                    if (lane + 32 < dr.nm) cnt2 = max(cnt2, (__float_as_int(W.man[(lane + 32) * MF_VEC4].x) >> 16) & 0xff);
                    two_step = dr.nm > 64 || __any_sync(ln.m, cnt2 > 1);
                }
            } else {
                // packed groups keep the list the schedule was built for in shared memory
                bool eq = cache.nm == dr.nm;
                for (int sl = lane; sl < dr.nm; sl += G)
                    eq = eq && W.sig[sl] == (uint16_t)(__float_as_int(W.man[sl * MF_VEC4].x) & 0xffff);
                same = __all_sync(ln.m, eq);
            }
            if (!same) {
                build_schedule(sc, W, ln, dr.nm, cache);
                cache.nm = dr.nm;
                cache.sig0 = sig0;
                cache.sig1 = sig1;
                if (G != 32) {
                    for (int sl = lane; sl < dr.nm; sl += G)
                        W.sig[sl] = (uint16_t)(__float_as_int(W.man[sl * MF_VEC4].x) & 0xffff);
                    __syncwarp(ln.m);
                }
            }
            const int rounds = cache.rounds;
            if (rounds >= 0) st.colour_passes += (unsigned)(rounds * sc.vel_iters);
            PH_T(t2);
            PH_ADD(2, t1, t2);
            const int iters = sc.vel_iters;
            auto pass = [&](unsigned e) {
                const int slr = e ? (int)(e & 0xffu) : 0xff;
                const int jr = e ? (int)((e >> 8) & 0x3fu) : lane;
                solve_pass<true>(W, bd, ln, slr, jr, (e & 0x8000u) != 0u);
            };
            auto pass1 = [&](unsigned e) {  // no manifold of the step has a second point
                const int slr = e ? (int)(e & 0xffu) : 0xff;
                const int jr = e ? (int)((e >> 8) & 0x3fu) : lane;
                solve_pass<false>(W, bd, ln, slr, jr, (e & 0x8000u) != 0u);
            };
            if (rounds < 0) {
                general_sweep<G>(sc, W, bd, ln, dr.nm, st);
            } else if (RR > 0 && rounds <= (RR < 2 ? RR : 2)) {
                // One or two rounds (most contact steps): the rows of this lane's manifolds stay in registers
                // for the whole sweep and so do the accumulated impulses, which start at zero every step and
                // are not read after the sweep - a pass is three shuffles and arithmetic, no shared memory.
                const unsigned ea = cache.ent[0], eb = (RR > 1 && rounds > 1) ? cache.ent[1] : 0u;
                const bool act_a = ea != 0u, act_b = eb != 0u;
                const bool st_a = (ea & 0x8000u) != 0u, st_b = (eb & 0x8000u) != 0u;
                const int j_a = act_a ? (int)((ea >> 8) & 0x3fu) : lane, j_b = act_b ? (int)((eb >> 8) & 0x3fu) : lane;
                const bool isa_a = st_a || lane < j_a, isa_b = st_b || lane < j_b;
                const MRows Ra = load_rows(W.man + (act_a ? (int)(ea & 0xffu) : 0) * MF_VEC4);
                float4 acc_a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                // sweeps whose manifolds all have a single point (84 % of the contact steps of cfg2) leave the second
                // point's rows out; not in packed launches, where the groups of a warp would take different variants
                // and the warp would run both (measured: -10 % at B = 7, K = 65536)
                const bool two_a = act_a && ((__float_as_int(Ra.m0.x) >> 16) & 0xff) > 1;
                if (RR < 2 || rounds == 1) {
                    const bool two = G < 32 || __any_sync(ln.m, two_a);
                    PH_T(tl0);
                    if (two) {
#pragma unroll kSweepUnroll
                        for (int it = 0; it < iters; ++it) solve_core<true>(Ra, acc_a, bd, ln, act_a, j_a, st_a, isa_a);
                    } else {
#pragma unroll kSweepUnroll
                        for (int it = 0; it < iters; ++it) solve_core<false>(Ra, acc_a, bd, ln, act_a, j_a, st_a, isa_a);
                    }
                    PH_T(tl1);
                    PH_ADD(10, tl0, tl1);  // the passes alone of one-round sweeps
                } else {
                    const MRows Rb = load_rows(W.man + (act_b ? (int)(eb & 0xffu) : 0) * MF_VEC4);
                    float4 acc_b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    const bool two_b = act_b && ((__float_as_int(Rb.m0.x) >> 16) & 0xff) > 1;
                    const bool two = G < 32 || __any_sync(ln.m, two_a || two_b);
                    if (two) {
#pragma unroll kSweepUnroll
                        for (int it = 0; it < iters; ++it) {
                            solve_core<true>(Ra, acc_a, bd, ln, act_a, j_a, st_a, isa_a);
                            solve_core<true>(Rb, acc_b, bd, ln, act_b, j_b, st_b, isa_b);
                        }
                    } else {
#pragma unroll kSweepUnroll
                        for (int it = 0; it < iters; ++it) {
                            solve_core<false>(Ra, acc_a, bd, ln, act_a, j_a, st_a, isa_a);
                            solve_core<false>(Rb, acc_b, bd, ln, act_b, j_b, st_b, isa_b);
                        }
                    }
                }
            } else if (RR >= 3 && rounds == 3) {
                // (build with the whole register file to itself, one CTA per SM: three and four rounds in registers too)
                RoundRegs r[3] = {load_round(W, cache.ent[0], lane), load_round(W, cache.ent[1], lane),
                                  load_round(W, cache.ent[2], lane)};
                if (__any_sync(ln.m, r[0].two || r[1].two || r[2].two))
                    sweep_in_registers<true, 3>(r, bd, ln, iters);
                else
                    sweep_in_registers<false, 3>(r, bd, ln, iters);
            } else if (RR >= 4 && rounds == 4) {
                RoundRegs r[4] = {load_round(W, cache.ent[0], lane), load_round(W, cache.ent[1], lane),
                                  load_round(W, cache.ent[2], lane), load_round(W, cache.ent[3], lane)};
                if (__any_sync(ln.m, r[0].two || r[1].two || r[2].two || r[3].two))
                    sweep_in_registers<true, 4>(r, bd, ln, iters);
                else
                    sweep_in_registers<false, 4>(r, bd, ln, iters);
            } else if (ROLLED && rounds <= 4) {
                // Launches that fill the machine are bound by instruction fetch (the step's code is larger than the
                // instruction cache and 16 warps per SM sit in different places of it): ONE copy of each pass variant in
                // a rolled loop over the rounds instead of four inlined copies.  Measured on B200 in the unpacked
                // 128-register build (profiles/r02_rolled_sweeps.txt): contact-rich launches gain (9 bodies, K = 40 960:
                // 5.18 -> 6.06 M rollouts/s; chains of two controls 3.76 -> 4.68; 15 bodies 5.98 -> 7.45), launches with
                // few contacts pay the loop's selects and branch (11 bodies 12.3 -> 11.5, 17 bodies 8.53 -> 8.24).
                const unsigned e0 = cache.ent[0], e1 = cache.ent[1], e2 = cache.ent[2], e3 = cache.ent[3];
                if (!two_step) {
                    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                        for (int r = 0; r < rounds; ++r) pass1(r == 0 ? e0 : r == 1 ? e1 : r == 2 ? e2 : e3);
                    }
                } else {
                    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
                        for (int r = 0; r < rounds; ++r) pass(r == 0 ? e0 : r == 1 ? e1 : r == 2 ? e2 : e3);
                    }
                }
            } else if (RR < 2 && rounds <= 4 && !two_step) {
                // (not in the build that keeps two rounds in registers: there this path only serves three and four
                // rounds, and a second copy of it cost more in instruction fetch than it saved - cfg2 2.05 -> 2.09 ms)
                for (int it = 0; it < iters; ++it) {
                    pass1(cache.ent[0]);
                    if (rounds > 1) pass1(cache.ent[1]);
                    if (rounds > 2) pass1(cache.ent[2]);
                    if (rounds > 3) pass1(cache.ent[3]);
                }
            } else if (rounds <= 4) {
                for (int it = 0; it < iters; ++it) {
                    pass(cache.ent[0]);
                    if (rounds > 1) pass(cache.ent[1]);
                    if (rounds > 2) pass(cache.ent[2]);
                    if (rounds > 3) pass(cache.ent[3]);
                }
            } else {
                for (int it = 0; it < iters; ++it) {
                    for (int r = 1; r <= rounds; ++r) {
                        unsigned e = 0u;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if ((int)(cache.ent[c] >> 16) == r) e = cache.ent[c];
                        pass(e);
                    }
                }
            }
            PH_T(t3);
            PH_ADD(3, t2, t3);
#ifdef MPS_PHASE_PROF
            st.ph[5] += 1;
            st.ph[9] += rounds < 0 ? 0 : rounds * iters;  // solver passes of the step
#endif
        }
    }
    if (bd.live) {
        const float ox = bd.x, oy = bd.y;
        bd.x = fmaf(sc.dt, bd.vx, bd.x);
        bd.y = fmaf(sc.dt, bd.vy, bd.y);
        bd.th = mps_wrap_angle(fmaf(sc.dt, bd.w, bd.th));
        mps_sincos(bd.th, &bd.s, &bd.c);
        // no point of the body moved further than |dx| + |dy| + R |dtheta| (slack for rounding)
        bd.mv += (fabsf(bd.x - ox) + fabsf(bd.y - oy) + bd.brad * fabsf(sc.dt * bd.w)) * 1.001f + 1e-7f;
    }
    publish_pose(W, bd, ln);
    PH_T(t4);
    PH_ADD(4, t1, t4);  // includes phases 2 and 3
}

__device__ __forceinline__ bool at_rest(const DevScene& sc, const Body& bd, const Ln ln)
{
    bool moving = false;
    if (bd.live) moving = (fmaf(bd.vx, bd.vx, bd.vy * bd.vy) > sc.v_rest2) || (fabsf(bd.w) > sc.w_rest);
    return !__any_sync(ln.m, moving);
}

// satisfiesBounds of the reference's compound space, per body in double
__device__ __forceinline__ bool body_in_bounds(const DevScene& sc, float xf, float yf, float thf)
{
    double x = (double)xf, y = (double)yf, th = (double)thf;
    if (x - MPS_DBL_EPS > (double)sc.x_max || x + MPS_DBL_EPS < (double)sc.x_min) return false;
    if (y - MPS_DBL_EPS > (double)sc.y_max || y + MPS_DBL_EPS < (double)sc.y_min) return false;
    if (!(th < MPS_PI_D + MPS_DBL_EPS && th > -MPS_PI_D - MPS_DBL_EPS)) return false;
    return true;
}

// isValid(state) with the state already in the lanes (sx, sy, sth = the float state values)
template <int G>
__device__ __forceinline__ uint32_t state_validity_flags(const DevScene& sc, const WarpMem& W, Body& bd, const Ln ln,
                                                         float sx, float sy, float sth, StepStats* st,
                                                         StepCache& cache)
{
    bool oob = bd.live && !body_in_bounds(sc, sx, sy, sth);
    if (__any_sync(ln.m, oob)) return MPS_F_BOUNDS;
    set_state(W, bd, ln, sx, sy, sth);
    DetectResult dr = detect<G>(sc, W, ln, false, cache);
    (void)st;
    uint32_t f = 0;
    if (dr.max_pen > sc.pen_max) f |= MPS_F_INFEASIBLE;
    if (dr.bad) f |= MPS_F_BAD_CONTACT;
    if (dr.overflow) f |= MPS_F_OVERFLOW;
    return f;
}

// sum_i mask_i w_i (pw |dxy_i| + ow arc_i) accumulated in body order, in double
__device__ __forceinline__ double warp_distance(const DevScene& sc, const Body& bd, const Ln ln, float sx, float sy,
                                                float sth, const float* __restrict__ other,
                                                const float* __restrict__ weights, uint32_t mask)
{
    const int lane = ln.g;
    double term = 0.0;
    if (bd.live && ((mask >> lane) & 1u)) {
        double cfg = mps_body_distance(sx, sy, sth, other[3 * lane], other[3 * lane + 1], other[3 * lane + 2],
                                       sc.pw, sc.ow);
        double w = weights ? (double)weights[lane] : 1.0;
        term = w * cfg;
    }
    double dist = 0.0;
    for (int i = 0; i < sc.B; ++i) {
        double t = __shfl_sync(ln.m, term, ln.abs_lane(i));
        if ((mask >> i) & 1u) dist += t;
    }
    return dist;
}

// MINB = CTAs per SM the register allocation is held to, G = lanes per chain.  Built: <4, 1, 32> (248 registers) for
// launches of one CTA per SM (K <= 3 chains per scheduler, cfg 2's extend): sweeps of up to four rounds keep their
// rows in registers (cfg2 1.333 -> 1.301 ms); <4, 3, 32> (up to 168
// registers) for launches with few chains, where one warp per scheduler runs alone and only its own
// dependent-instruction latency matters; <4, 4, 32> (128 registers) for launches that fill the machine;
// <4, 4, 8> and <4, 4, 16> with four / two chains per warp (see mps_launch_rollout).  None spills.
template <int WARPS, int MINB, int G, bool ROLLED = false>
__global__ void __launch_bounds__(WARPS * 32, MINB) rollout_kernel(RolloutParams p)
{
    MPS_DYN_SMEM(smem_dyn);
    // The base of the dynamic shared memory goes through an opaque register once: otherwise the compiler
    // rematerialises the symbol's address at every use, each time with a read of a special register.
    unsigned char* smem_raw = opaque_shared_base(smem_dyn);
    DevScene& sc = *reinterpret_cast<DevScene*>(smem_raw);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.scene);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
        for (int i = threadIdx.x; i < (int)(sizeof(DevScene) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (p.active_warps > 0 && (int)(threadIdx.x >> 5) >= p.active_warps) return;
    const Ln ln = make_ln<G>();
    const int lane = ln.g;  // lane within the chain's group = body index
    const int group = (threadIdx.x >> 5) * (32 / G) + ((threadIdx.x & 31) / G);
    const int B = sc.B;
    // a packed launch may give every group fewer manifold slots than the scene allows: a chain that needs more
    // is left to the one-chain-per-warp kernel that follows (p.redo)
    const bool small_cap = G < 32 && p.mcap_small > 0;
    const int Mcap = small_cap ? p.mcap_small : sc.Mcap;
    unsigned char* wbase = smem_raw + p.scene_bytes + (size_t)group * p.warp_bytes;
    const WarpMem W = carve_warp_mem(sc, wbase, Mcap);
    publish_statics(sc, W, ln, G);

    Body bd;
    bd.live = lane < B;
    bd.dynamic = bd.live && lane != sc.robot;
    {
        int li = bd.live ? lane : 0;
        bd.inv_m = bd.live ? sc.inv_m[li] : 0.0f;
        bd.inv_i = bd.live ? sc.inv_i[li] : 0.0f;
        bd.mass = sc.mass[li];
        bd.inertia = sc.inertia[li];
        bd.lin_cap = sc.lin_cap[li];
        bd.brad = sc.brad[li];
        bd.ang_cap = sc.ang_cap[li];
    }
    bd.x = bd.y = bd.th = 0.0f;
    bd.c = 1.0f;
    bd.s = 0.0f;
    bd.vx = bd.vy = bd.w = 0.0f;

    StepStats st = {0u, 0u, 0u, 0u};
    StepCache cache;
    cache.reset();
    const int L = p.L;

    for (;;) {
        int k = 0;
        if (lane == 0) k = (int)atomicAdd(p.work_counter, 1u);
        k = __shfl_sync(ln.m, k, ln.abs_lane(0));
        if (k >= p.K) break;
        if (p.order) k = p.order[k];  // longest chains first (mps_launch_chain_order)
        if (p.redo_only && !p.redo[k]) continue;
        bool redo = false;
        int n = p.n_ctrl ? p.n_ctrl[k] : L;
        if (n > L) n = L;
        const long long t_begin = clock64();
        cache.reset();
        bd.mv = 0.0f;
#ifdef MPS_PHASE_PROF
        for (int i = 0; i < 11; ++i) st.ph[i] = 0;
#endif
        // start state of the chain
        float sx = 0.0f, sy = 0.0f, sth = 0.0f;
        const float* dest = p.dest;
        uint32_t dmask = p.mask;
        if (p.K_per > 0) {
            // multi-instance batch: start = a node of this instance's forest segment
            const int m = k / p.K_per;
            const size_t col = (size_t)p.inst_ids[m] * p.f_seg + (size_t)p.parents[m];
            if (bd.live) {
                sx = p.f_states[(size_t)(3 * lane) * p.f_cols + col];
                sy = p.f_states[(size_t)(3 * lane + 1) * p.f_cols + col];
                sth = p.f_states[(size_t)(3 * lane + 2) * p.f_cols + col];
            }
            if (dest) dest += (size_t)m * 3 * B;
            if (p.masks) dmask = p.masks[m];
        } else if (bd.live) {
            const float* st0 = p.start + (size_t)k * p.start_stride;
            sx = st0[3 * lane];
            sy = st0[3 * lane + 1];
            sth = st0[3 * lane + 2];
        }
        int n_ok = 0, goal_step = -1;
        int l = 0;
        for (; l < n; ++l) {
            const size_t kl = (size_t)k * L + l;
            const float* ctrl = p.controls + kl * MPS_CONTROL_DIM;
            uint32_t flags = MPS_F_RAN;
            unsigned steps0 = st.steps;
            set_state(W, bd, ln, sx, sy, sth);
            // RampVelocityControl::computeRamp (rest time reset to 0 before propagating)
            float v0 = ctrl[0], v1 = ctrl[1], v2 = ctrl[2], t_pl = ctrl[3];
            float t_acc = fabsf(v0 / sc.accel[0]);
            {
                float t = fabsf(v0 / sc.accel[0]);
                t_acc = t > t_acc ? t : t_acc;
                t = fabsf(v1 / sc.accel[1]);
                t_acc = t > t_acc ? t : t_acc;
                t = fabsf(v2 / sc.accel[2]);
                t_acc = t > t_acc ? t : t_acc;
            }
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
            if (t_acc != 0.0f) {
                float inv = 1.0f / t_acc;
                a0 = inv * v0;
                a1 = inv * v1;
                a2 = inv * v2;
            }
            const float T = 2.0f * t_acc + t_pl + 0.0f;
            const float t_ap = t_acc + t_pl;
            const double t_dec_end = 2.0 * (double)t_acc + (double)t_pl;
            float time = 0.0f;
            bool ok = true;
            // Active phase (SimEnvStatePropagator.cpp:75-83) and rest phase (:86-105) run through ONE loop with
            // one call of physics_step: the step is the bulk of the kernel's code, and a second inlined copy
            // costs more in instruction-cache misses than the phase test costs per step.
            float rest = 0.0f;
            bool in_rest = false, resting = false;
            int jam = 0;
            for (;;) {
                if (!in_rest && !(time < T && ok)) {
                    if (!ok) break;
                    in_rest = true;
                    resting = at_rest(sc, bd, ln);
                }
                if (in_rest && !(rest <= sc.t_max && !resting && ok)) break;
                float ux = 0.0f, uy = 0.0f, uw = 0.0f;
                if (!in_rest) {
                    if (time < t_acc) {
                        ux = time * a0;
                        uy = time * a1;
                        uw = time * a2;
                    } else if (time < t_ap) {
                        ux = v0;
                        uy = v1;
                        uw = v2;
                    } else if ((double)time < t_dec_end) {
                        float s = time - t_acc - t_pl;
                        ux = v0 - s * a0;
                        uy = v1 - s * a1;
                        uw = v2 - s * a2;
                    }
                }
                bool bad, over;
                float pen;
                physics_step<G, (MINB == 1 ? MPS_REG_ROUNDS_SOLO : MINB <= 3 ? 2 : (G < 32 ? MPS_REG_ROUNDS : 0)), ROLLED>(sc, W, bd, ln, ux, uy, uw, bad, over, pen, st, cache);
                if (over && small_cap) {
                    redo = true;
                    break;
                }
                if (over) {
                    flags |= MPS_F_OVERFLOW;
                    ok = false;
                }
                if (!in_rest) {
                    time += sc.dt;
                    if (bad && sc.strict) {
                        flags |= MPS_F_ACTIVE_CONTACT;
                        ok = false;
                    }
                } else {
                    if (bad) {
                        flags |= MPS_F_REST_CONTACT;
                        ok = false;
                    }
                    rest += sc.dt;
                    // jam rule: wedged deeper than 1.25 * slop for jam_steps consecutive rest steps
                    jam = pen > 1.25f * sc.slop ? jam + 1 : 0;
                    if (sc.jam_steps > 0 && jam >= sc.jam_steps) {
                        flags |= MPS_F_JAMMED;
                        ok = false;
                    }
                    resting = at_rest(sc, bd, ln);
                }
            }
            if (in_rest) {
                if (!(rest <= sc.t_max && resting)) flags |= MPS_F_NOT_AT_REST;
                ok = rest <= sc.t_max && resting && ok;
            }
            if (redo) break;
            // extractState: theta normalised in double, read back as float
            float rx = bd.x, ry = bd.y;
            float rth = (float)mps_normalize_orientation((double)bd.th);
            uint32_t vf = state_validity_flags<G>(sc, W, bd, ln, rx, ry, rth, &st, cache);
            if ((vf & MPS_F_OVERFLOW) && small_cap) {
                redo = true;
                break;
            }
            flags |= vf;
            ok = ok && vf == 0u;
            if (ok) {
                flags |= MPS_F_OK;
                // goal test on the result (RRT.cpp:377, :484, :726)
                if (p.n_goals > 0) {
                    double gd = 0.0;
                    for (int g = 0; g < p.n_goals; ++g) {
                        int gb = p.goals[g].body;
                        float gx = __shfl_sync(ln.m, rx, ln.abs_lane(gb));
                        float gy = __shfl_sync(ln.m, ry, ln.abs_lane(gb));
                        float fx = p.goals[g].x - gx;
                        float fy = p.goals[g].y - gy;
                        double dc = sqrt((double)fx * (double)fx + (double)fy * (double)fy);
                        double e = dc - (double)p.goals[g].tol;
                        gd += e > 0.0 ? e : 0.0;
                    }
                    if (gd <= MPS_DBL_EPS) {
                        flags |= MPS_F_GOAL;
                        if (goal_step < 0) goal_step = l;
                    }
                }
                if (p.ball_center) {
                    uint32_t arr_mask = ~(1u << sc.robot);
                    if (B < 32) arr_mask &= (1u << B) - 1u;
                    double bdist = warp_distance(sc, bd, ln, rx, ry, rth, p.ball_center, p.weights, arr_mask);
                    if (bdist <= (double)p.ball_radius) flags |= MPS_F_IN_BALL;
                }
            }
            if (bd.live) {
                float* os = p.out_states + kl * 3 * B + 3 * lane;
                os[0] = rx;
                os[1] = ry;
                os[2] = rth;
            }
            if (lane == 0) {
                p.out_rest[kl] = rest;
                p.out_flags[kl] = flags;
                p.out_steps[kl] = (int)(st.steps - steps0);
            }
            if (!ok) {
                ++l;
                break;
            }
            ++n_ok;
            sx = rx;
            sy = ry;
            sth = rth;
            if (p.stop_at_goal && goal_step >= 0) {
                ++l;
                break;
            }
        }
        if (redo) {
            if (lane == 0) p.redo[k] = 1;
            continue;
        }
        // controls of the chain that were never propagated
        for (; l < L; ++l) {
            const size_t kl = (size_t)k * L + l;
            if (bd.live) {
                float* os = p.out_states + kl * 3 * B + 3 * lane;
                os[0] = 0.0f;
                os[1] = 0.0f;
                os[2] = 0.0f;
            }
            if (lane == 0) {
                p.out_rest[kl] = p.controls[kl * MPS_CONTROL_DIM + 4];
                p.out_flags[kl] = 0u;
                p.out_steps[kl] = 0;
            }
        }
        double dist = __longlong_as_double(0x7ff0000000000000ll);  // +inf
        if (n_ok > 0 && dest) dist = warp_distance(sc, bd, ln, sx, sy, sth, dest, p.weights, dmask);
        if (lane == 0) {
            p.out_dist[k] = dist;
            p.out_nok[k] = n_ok;
            p.out_goal_step[k] = goal_step;
#ifdef MPS_PHASE_PROF
            if (p.out_cycles) {
                p.out_cycles[(size_t)k * 12] = clock64() - t_begin;
                for (int i = 0; i < 11; ++i) p.out_cycles[(size_t)k * 12 + 1 + i] = st.ph[i];
            }
#else
            if (p.out_cycles) p.out_cycles[k] = clock64() - t_begin;
#endif
        }
    }
    // counters for the roofline accounting
    if (lane == 0 && p.counters) {
        atomicAdd(&p.counters[0], (unsigned long long)st.steps);
        atomicAdd(&p.counters[1], (unsigned long long)st.candidates);
        atomicAdd(&p.counters[2], (unsigned long long)st.touching);
        atomicAdd(&p.counters[3], (unsigned long long)st.colour_passes);
    }
}

// Order in which the warps pull the chains: longest expected first.  A launch lasts as long as its last chain, and
// chain lengths vary by 10x; the number of active-phase steps of a chain is known from its controls (ramp
// durations), the rest phase adds a few dozen per control.  Chains are bucketed by that estimate (8 steps per
// bucket, counting sort by one CTA); the order inside a bucket is arbitrary - no result depends on the order.
__global__ void __launch_bounds__(1024) chain_order_kernel(const float* __restrict__ controls,
                                                           const int32_t* __restrict__ n_ctrl, int K, int L, float inv_ax,
                                                           float inv_ay, float inv_aw, float inv_dt, int32_t* __restrict__ order)
{
    __shared__ unsigned int hist[256];
    __shared__ unsigned int offs[256];
    const int tid = threadIdx.x;
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    auto bucket_of = [&](int k) -> int {
        int n = n_ctrl ? n_ctrl[k] : L;
        n = n > L ? L : n;
        float steps = 0.0f;
        for (int l = 0; l < n; ++l) {
            const float* c = controls + ((size_t)k * L + l) * MPS_CONTROL_DIM;
            const float t_acc = fmaxf(fmaxf(fabsf(c[0] * inv_ax), fabsf(c[1] * inv_ay)), fabsf(c[2] * inv_aw));
            steps += (2.0f * t_acc + c[3]) * inv_dt + 32.0f;
        }
        int bkt = (int)(steps * 0.125f);
        bkt = bkt < 0 ? 0 : (bkt > 255 ? 255 : bkt);
        return 255 - bkt;  // bucket 0 = the longest chains
    };
    for (int k = tid; k < K; k += blockDim.x) atomicAdd(&hist[bucket_of(k)], 1u);
    __syncthreads();
    if (tid < 32) {
        // exclusive scan of the 256 counts by one warp, 8 per lane
        unsigned int loc[8], sum = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            loc[i] = sum;
            sum += hist[tid * 8 + i];
        }
        unsigned int inc = sum;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned int o = __shfl_up_sync(FULL, inc, off);
            if (tid >= off) inc += o;
        }
        const unsigned int base = inc - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) offs[tid * 8 + i] = base + loc[i];
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x) order[atomicAdd(&offs[bucket_of(k)], 1u)] = k;
}

// argmin over the K chains: strict <, first index wins (NaiveControlSampler.cpp:55, RRT.cpp:467);
// copies the winner chain into the best buffers and optionally appends it to the tree.
__global__ void __launch_bounds__(1024) select_kernel(SelectParams p)
{
    __shared__ double s_key[32];
    __shared__ int s_idx[32];
    const int tid = threadIdx.x;
    const int m = blockIdx.x;  // instance
    const int k_base = m * p.K;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double best = inf;
    int best_k = -1;
    for (int k = k_base + tid; k < k_base + p.K; k += blockDim.x) {
        if (p.nok[k] <= 0) continue;
        double d = p.dist[k];
        double key = !p.has_dest ? 0.0 : (p.compare_f32 ? (double)(float)d : d);
        // the reference starts from max() and compares with <: a key that is not below it never wins
        double limit = p.compare_f32 ? (double)3.402823466e+38f : 1.7976931348623157e+308;
        if (p.has_dest && !(key < limit)) continue;
        if (best_k < 0 || key < best) {
            best = key;
            best_k = k;
        }
    }
    // warp then block reduction on (key, k): smaller key, then smaller k
    for (int off = 16; off > 0; off >>= 1) {
        double ob = __shfl_down_sync(FULL, best, off);
        int ok = __shfl_down_sync(FULL, best_k, off);
        if (ok >= 0 && (best_k < 0 || ob < best || (ob == best && ok < best_k))) {
            best = ob;
            best_k = ok;
        }
    }
    if ((tid & 31) == 0) {
        s_key[tid >> 5] = best;
        s_idx[tid >> 5] = best_k;
    }
    __syncthreads();
    if (tid < 32) {
        int nw = (blockDim.x + 31) >> 5;
        best = tid < nw ? s_key[tid] : inf;
        best_k = tid < nw ? s_idx[tid] : -1;
        for (int off = 16; off > 0; off >>= 1) {
            double ob = __shfl_down_sync(FULL, best, off);
            int ok = __shfl_down_sync(FULL, best_k, off);
            if (ok >= 0 && (best_k < 0 || ob < best || (ob == best && ok < best_k))) {
                best = ob;
                best_k = ok;
            }
        }
        if (tid == 0) {
            s_idx[0] = best_k;
        }
    }
    __syncthreads();
    best_k = s_idx[0];
    const int L = p.L, B = p.B;
    int n_ok = best_k >= 0 ? p.nok[best_k] : 0;
    // nodes appended: the chain up to and including the first goal state (HybridActionRRT::extend stops
    // adding there, RRT.cpp:479-488), else the whole successful prefix
    int n_app = n_ok;
    if (best_k >= 0 && p.goal_step[best_k] >= 0) n_app = p.goal_step[best_k] + 1;
    const int seg_id = p.inst_ids ? p.inst_ids[m] : 0;
    int first = -1;
    if (best_k >= 0 && p.append && tid == 0) {
        // tree append of the winner chain (RearrangementRRT::addToTree, RRT.cpp:234-244)
        long long n0 = p.tree_size[seg_id];
        if ((size_t)(n0 + n_app) <= p.tree_seg) {
            first = (int)n0;
            p.tree_size[seg_id] = n0 + n_app;
        } else {
            first = -2;  // segment full: nothing appended
        }
    }
    mps_extend_best* rec = p.best + m;
    if (tid == 0) {
        rec->k = best_k < 0 ? -1 : best_k - k_base + p.k_offset;
        rec->n_ok = n_ok;
        rec->goal_step = best_k < 0 ? -1 : p.goal_step[best_k];
        rec->distance = best_k < 0 ? inf : p.dist[best_k];
        rec->tree_index = (best_k >= 0 && p.append) ? first : -1;
        s_idx[1] = first;
    }
    __syncthreads();
    first = s_idx[1];
    float* bs = p.best_states + (size_t)m * L * 3 * B;
    float* bc = p.best_controls + (size_t)m * L * MPS_CONTROL_DIM;
    uint32_t* bf = p.best_flags + (size_t)m * L;
    for (int i = tid; i < L * 3 * B; i += blockDim.x)
        bs[i] = best_k < 0 ? 0.0f : p.states[(size_t)best_k * L * 3 * B + i];
    for (int i = tid; i < L * MPS_CONTROL_DIM; i += blockDim.x) {
        float v = 0.0f;
        if (best_k >= 0) {
            int l = i / MPS_CONTROL_DIM, c = i % MPS_CONTROL_DIM;
            v = c == 4 ? p.rest[(size_t)best_k * L + l] : p.controls[((size_t)best_k * L + l) * MPS_CONTROL_DIM + c];
        }
        bc[i] = v;
    }
    for (int i = tid; i < L; i += blockDim.x) bf[i] = best_k < 0 ? 0u : p.flags[(size_t)best_k * L + i];
    if (best_k >= 0 && p.append && first >= 0) {
        const size_t cap = p.tree_cap;
        const size_t col0 = (size_t)seg_id * p.tree_seg + (size_t)first;
        const int parent = p.parents ? p.parents[m] : p.parent;
        for (int i = tid; i < n_app * 3 * B; i += blockDim.x) {
            int l = i / (3 * B), r = i % (3 * B);
            p.tree_states[(size_t)r * cap + col0 + l] = p.states[((size_t)best_k * L + l) * 3 * B + r];
        }
        for (int i = tid; i < n_app * MPS_CONTROL_DIM; i += blockDim.x) {
            int l = i / MPS_CONTROL_DIM, c = i % MPS_CONTROL_DIM;
            float v = c == 4 ? p.rest[(size_t)best_k * L + l]
                             : p.controls[((size_t)best_k * L + l) * MPS_CONTROL_DIM + c];
            p.tree_controls[(size_t)c * cap + col0 + l] = v;
        }
        for (int i = tid; i < n_app; i += blockDim.x) {
            p.tree_parent[col0 + i] = i == 0 ? parent : first + i - 1;
            p.tree_slice[col0 + i] = -1;
        }
    }
    if (p.done_flag) {  // single instance: this CTA is the last work of the extend
        __threadfence_system();
        __syncthreads();
        if (tid == 0) *p.done_flag = p.done_seq;
    }
}

// Exchange step of the K-sharded extend: every rank holds the all-gathered winner records in rank
// order and applies the reference's rule to them - strict <, the first index wins (NaiveControlSampler.cpp:55,
// RRT.cpp:467; on equal keys the lower global chain index, which for contiguous shards is the lower rank) - so
// all ranks agree on the winner without another exchange.  The winner's record becomes this ctx's record; the
// GPU that owns the tree appends the winner chain (RearrangementRRT::addToTree RRT.cpp:234-244, up to the first
// goal state :479-488).  One CTA.
__global__ void __launch_bounds__(256) merge_records_kernel(MergeParams p)
{
    __shared__ int s_win, s_first;
    const int tid = threadIdx.x;
    const int L = p.L, B = p.B;
    if (tid == 0) {
        int win = -1;
        double best = 0.0;
        int best_k = 0;
        for (int r = 0; r < p.world; ++r) {
            const mps_extend_best* h = reinterpret_cast<const mps_extend_best*>(p.gathered + (size_t)r * p.rec_bytes);
            if (h->k < 0) continue;
            const double key = !p.has_dest ? 0.0 : (p.compare_f32 ? (double)(float)h->distance : h->distance);
            if (win < 0 || key < best || (key == best && h->k < best_k)) {
                win = r;
                best = key;
                best_k = h->k;
            }
        }
        s_win = win;
        int first = -1;
        if (win >= 0 && p.append) {
            const mps_extend_best* h = reinterpret_cast<const mps_extend_best*>(p.gathered + (size_t)win * p.rec_bytes);
            const int n_app = h->goal_step >= 0 ? h->goal_step + 1 : h->n_ok;
            const long long n0 = p.tree_size[0];
            if ((size_t)(n0 + n_app) <= p.tree_cap) {
                first = (int)n0;
                p.tree_size[0] = n0 + n_app;
            } else {
                first = -2;
            }
        }
        s_first = first;
    }
    __syncthreads();
    const int win = s_win, first = s_first;
    const unsigned char* src = p.gathered + (size_t)(win < 0 ? 0 : win) * p.rec_bytes;
    // the record travels as 32-bit words (it is 8-byte aligned and a multiple of 8 bytes long)
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
    uint32_t* dw = reinterpret_cast<uint32_t*>(p.record);
    for (size_t i = tid; i < p.rec_bytes / 4; i += blockDim.x) dw[i] = sw[i];
    __syncthreads();
    if (tid == 0) {
        mps_extend_best* h = reinterpret_cast<mps_extend_best*>(p.record);
        if (win < 0) {
            h->k = -1;
            h->n_ok = 0;
            h->goal_step = -1;
            h->distance = __longlong_as_double(0x7ff0000000000000ll);
        }
        h->tree_index = (win >= 0 && p.append) ? first : -1;
        if (p.head_out) *p.head_out = *h;
    }
    if (win >= 0 && p.append && first >= 0) {
        const mps_extend_best* h = reinterpret_cast<const mps_extend_best*>(src);
        const int n_app = h->goal_step >= 0 ? h->goal_step + 1 : h->n_ok;
        const float* st = reinterpret_cast<const float*>(src + sizeof(mps_extend_best));
        const float* ct = st + (size_t)L * 3 * B;
        const size_t cap = p.tree_cap;
        for (int i = tid; i < n_app * 3 * B; i += blockDim.x) {
            const int l = i / (3 * B), r = i % (3 * B);
            p.tree_states[(size_t)r * cap + first + l] = st[(size_t)l * 3 * B + r];
        }
        for (int i = tid; i < n_app * MPS_CONTROL_DIM; i += blockDim.x) {
            const int l = i / MPS_CONTROL_DIM, c = i % MPS_CONTROL_DIM;
            p.tree_controls[(size_t)c * cap + first + l] = ct[(size_t)l * MPS_CONTROL_DIM + c];
        }
        for (int i = tid; i < n_app; i += blockDim.x) {
            p.tree_parent[first + i] = i == 0 ? p.parent : first + i - 1;
            p.tree_slice[first + i] = -1;
        }
    }
}

// isValid for M states, one warp per state
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) validity_kernel(ValidityParams p)
{
    MPS_DYN_SMEM(smem_dyn);
    // The base of the dynamic shared memory goes through an opaque register once: otherwise the compiler
    // rematerialises the symbol's address at every use, each time with a read of a special register.
    unsigned char* smem_raw = opaque_shared_base(smem_dyn);
    DevScene& sc = *reinterpret_cast<DevScene*>(smem_raw);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.scene);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
        for (int i = threadIdx.x; i < (int)(sizeof(DevScene) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int B = sc.B;
    unsigned char* wbase = smem_raw + p.scene_bytes + (size_t)warp * p.warp_bytes;
    const WarpMem W = carve_warp_mem(sc, wbase, sc.Mcap);
    const Ln ln = make_ln<32>();
    publish_statics(sc, W, ln, 32);
    Body bd;
    bd.live = lane < B;
    bd.dynamic = false;
    bd.inv_m = bd.inv_i = bd.mass = bd.inertia = bd.lin_cap = bd.ang_cap = 0.0f;
    bd.mv = 0.0f;
    bd.brad = 0.0f;
    bd.x = bd.y = bd.th = 0.0f;
    bd.c = 1.0f;
    bd.s = 0.0f;
    bd.vx = bd.vy = bd.w = 0.0f;
    for (int i = blockIdx.x * WARPS + warp; i < p.n; i += gridDim.x * WARPS) {
        float sx = 0.0f, sy = 0.0f, sth = 0.0f;
        if (bd.live) {
            sx = p.states[(size_t)i * 3 * B + 3 * lane];
            sy = p.states[(size_t)i * 3 * B + 3 * lane + 1];
            sth = p.states[(size_t)i * 3 * B + 3 * lane + 2];
        }
        StepCache cache;
        cache.reset();
        uint32_t f = state_validity_flags<32>(sc, W, bd, ln, sx, sy, sth, nullptr, cache);
        if (lane == 0) {
            p.flags[i] = f;
            p.valid[i] = f == 0u ? 1 : 0;
        }
    }
    if (p.done_flag) {  // single CTA (mps_launch_validity)
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *p.done_flag = p.done_seq;
    }
}

}  // namespace

size_t mps_rollout_warp_bytes(const DevScene& sc) { return mps_rollout_group_bytes(sc, sc.Mcap); }

// shared memory of one chain (one lane group) with `mcap` manifold slots
size_t mps_rollout_group_bytes(const DevScene& sc, int mcap)
{
    size_t b = 5 * sizeof(float) * (size_t)mps_pose_slots(sc.B, sc.S) + (size_t)MF_FIELDS * mcap * sizeof(float) +
               (((size_t)sc.B * sc.slot_stride + 3) & ~(size_t)3) + sizeof(uint16_t) * (size_t)((sc.P + 1) & ~1) +
               sizeof(uint16_t) * (size_t)mcap;
    return (b + 15) & ~(size_t)15;
}

size_t mps_rollout_scene_bytes() { return (sizeof(DevScene) + 15) & ~(size_t)15; }

#ifndef MPS_CPU_EMU  // the launchers need the CUDA runtime; the test-side emulator (tests/emu) drives the kernels itself
#define ROLLOUT_WARPS 4

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}


// grid of a rollout launch: as many CTAs as fit the machine, at most one group of lanes per chain
template <class Kernel>
static cudaError_t launch_rollout_kernel(Kernel kernel, std::atomic<bool>* attr_set, RolloutParams& p, int threads,
                                         int groups, int num_sms, cudaStream_t stream, int ctas_per_sm, int cap_per_sm)
{
    size_t smem = p.scene_bytes + (size_t)groups * p.warp_bytes;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {  // the attribute is per device: a process may hold contexts on several
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    int blocks_per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int want = (p.K + groups - 1) / groups;
    // tuning aid: MPS_ROLLOUT_CTAS_PER_SM caps the resident CTAs per SM
    if (cap_per_sm > 0)
        ctas_per_sm = cap_per_sm;
    if (ctas_per_sm > 0 && blocks_per_sm > ctas_per_sm) blocks_per_sm = ctas_per_sm;
    int grid = num_sms * blocks_per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    e = cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    kernel<<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

// WARPS warps per CTA, MINB CTAs per SM for the register budget, G lanes per chain
template <int WARPS, int MINB, int G, bool ROLLED = false>
static cudaError_t launch_rollout_t(RolloutParams& p, const RolloutTuning& tune, int num_sms, cudaStream_t stream,
                                    int ctas_per_sm = 0)
{
    if (G < 32 && p.mcap_small > 0) p.warp_bytes = p.warp_bytes_small;
    static std::atomic<bool> attr_set[64];
    return launch_rollout_kernel(rollout_kernel<WARPS, MINB, G, ROLLED>, attr_set, p, WARPS * 32, WARPS * (32 / G), num_sms, stream,
                                 ctas_per_sm, tune.ctas_per_sm);
}

RolloutTuning mps_rollout_tuning_from_env()
{
    RolloutTuning t;
    t.group = env_int("MPS_ROLLOUT_GROUP", 0);
    t.minb = env_int("MPS_ROLLOUT_MINB", 0);
    t.ctas_per_sm = env_int("MPS_ROLLOUT_CTAS_PER_SM", 0);
    t.active_warps = env_int("MPS_ROLLOUT_ACTIVE_WARPS", 0);
    t.solo_build = env_int("MPS_ROLLOUT_SOLO_BUILD", 1);
    t.rolled_max_bodies = env_int("MPS_ROLLED_MAX_BODIES", 12);
    return t;
}

cudaError_t mps_launch_rollout(const RolloutParams& p_in, const RolloutTuning& tune, int num_sms, int n_bodies,
                               cudaStream_t stream)
{
    RolloutParams p = p_in;
    // Which build and grid (measured on B200, profiles/tools/few_probe.py and packed_probe.py, cfg 2's scene; W =
    // chains per warp scheduler):
    //   W <= 3   one chain per warp, 168-register build, ONE CTA per SM.  The launch lasts as long as its slowest
    //            chain, and a chain is faster when its warp has a scheduler to itself; with the longest chains pulled
    //            first (mps_launch_chain_order) the short ones queue behind them without stretching the launch
    //            (K = 1024: 1.41 -> 1.35 ms against a grid that leaves two CTAs on 108 of the 148 SMs).
    //   W <= 8   the same build, two CTAs per SM (K = 4096 chains of <= 4 controls: 1.93 ms; 3.05 with one CTA per
    //            SM, 2.12 with three, 3.83 packed).
    //   beyond   128-register build, as many CTAs as fit (issue-bound regime).
    //   packed   four chains per warp (scenes with at most 8 bodies) pay off only for many single-control chains:
    //            K = 16 384: 1.00 ms against 1.26, K = 65 536: 3.39 against 4.76; K = 8192: 0.87 against 0.68, and
    //            chains of several controls (diverging groups) lose at every K measured (K = 16 384: 6.57 against 5.68).
    // Tuning aids: MPS_ROLLOUT_MINB = 3 | 4 pins the register budget of the unpacked build, MPS_ROLLOUT_GROUP =
    // 8 | 16 | 32 the lanes per chain, MPS_ROLLOUT_CTAS_PER_SM the resident CTAs.
    p.active_warps = tune.active_warps;
    const int forced_minb = tune.minb;
    const int forced_g = tune.group;
    const double per_sched = (double)p.K / ((double)num_sms * ROLLOUT_WARPS);
    int g = forced_g;
    // two chains per warp (9 .. 16 bodies): with the sweeps rolled over the rounds they gain 5 .. 12 % in single-control
    // launches of K >= 16 384 from ONE start state on scenes of 9 .. 12 bodies (profiles/r02_g16_rolled.txt) and lose
    // below that K, with chains of several controls, with 13 and more bodies, and in multi-instance launches (every
    // instance its own start state: the two chains of a warp share less of the instruction stream; config 5's sweep
    // 2.08 -> 2.28 s)
    if (g == 0)
        g = (n_bodies <= 8 && p.L == 1 && p.K >= 12288)                                      ? 8
            : (n_bodies >= 9 && n_bodies <= 12 && p.L == 1 && p.K >= 16384 && p.K_per == 0) ? 16
                                                                                              : 32;
    if (g == 8 && n_bodies > 8) g = 16;
    if (g == 16 && n_bodies > 16) g = 32;
    if (g != 32) {
        // packed launch with the small manifold capacity, then the chains it could not hold (normally none)
        // once more, one per warp with the full capacity
        cudaError_t e = cudaMemsetAsync(p.redo, 0, (size_t)p.K, stream);
        if (e != cudaSuccess) return e;
        RolloutParams q = p;
        q.redo_only = 0;
        // (two chains per warp always with the sweeps rolled over the rounds: 9 .. 15 bodies, +16 .. 28 % over the inlined
        // sweeps of the same packing, profiles/r02_g16_rolled.txt)
        e = g == 8 ? launch_rollout_t<4, 4, 8>(q, tune, num_sms, stream) : launch_rollout_t<ROLLOUT_WARPS, 4, 16, true>(q, tune, num_sms, stream);
        if (e != cudaSuccess) return e;
        if (p.mcap_small <= 0) return cudaSuccess;
        p.redo_only = 1;
        p.mcap_small = 0;
        return launch_rollout_t<ROLLOUT_WARPS, 4, 32>(p, tune, num_sms, stream);
    }
    p.redo_only = 0;
    p.mcap_small = 0;
    int minb = forced_minb;
    if (minb == 0) minb = per_sched <= 8.0 ? 3 : 4;
    int ctas = 0;
    if (p.order && per_sched <= 3.0)
        ctas = 1;
    else if (p.order && per_sched <= 8.0 && p.L > 1)
        ctas = 2;  // (single-control chains vary less in length: three CTAs per SM are better there, 0.41 against 0.51 ms at K = 4096)
    // one CTA per SM: the build that may use the whole register file (sweeps of up to four rounds in registers)
    const int solo_build = tune.solo_build;
    if (minb == 3 && ctas == 1 && solo_build) return launch_rollout_t<ROLLOUT_WARPS, 1, 32>(p, tune, num_sms, stream, ctas);
    if (minb == 3) return launch_rollout_t<ROLLOUT_WARPS, 3, 32>(p, tune, num_sms, stream, ctas);
    // launches that fill the machine: the build with the sweeps rolled over the rounds (see physics_step; measured over
    // seven shapes, profiles/r02_rolled_sweeps.txt: -6 .. +25 %, +7 % on average) for scenes of up to 12 bodies - the
    // 17-body clutter of config 4 loses 3 % with it, config 5's sweep (9 bodies) gains 4 %.  MPS_ROLLED_MAX_BODIES = 0 / 32
    // select one build for everything, for comparison.
    const int rolled_max = tune.rolled_max_bodies;
    if (n_bodies <= rolled_max) return launch_rollout_t<ROLLOUT_WARPS, 4, 32, true>(p, tune, num_sms, stream, ctas);
    return launch_rollout_t<ROLLOUT_WARPS, 4, 32>(p, tune, num_sms, stream, ctas);
}

cudaError_t mps_launch_chain_order(const float* controls, const int32_t* n_ctrl, int K, int L, const float accel[3],
                                   float dt, int32_t* order, cudaStream_t stream)
{
    chain_order_kernel<<<1, 1024, 0, stream>>>(controls, n_ctrl, K, L, 1.0f / accel[0], 1.0f / accel[1], 1.0f / accel[2],
                                               1.0f / dt, order);
    return cudaGetLastError();
}

cudaError_t mps_launch_select(const SelectParams& p, cudaStream_t stream)
{
    int threads = p.K >= 1024 ? 1024 : ((p.K + 31) / 32) * 32;
    if (threads < 32) threads = 32;
    select_kernel<<<p.M, threads, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t mps_launch_merge_records(const MergeParams& p, cudaStream_t stream)
{
    merge_records_kernel<<<1, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t mps_launch_validity(const ValidityParams& p, int num_sms, cudaStream_t stream)
{
    size_t smem = p.scene_bytes + (size_t)ROLLOUT_WARPS * p.warp_bytes;
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(validity_kernel<ROLLOUT_WARPS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    int want = (p.n + ROLLOUT_WARPS - 1) / ROLLOUT_WARPS;
    int grid = num_sms * 4;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    if (p.done_flag) grid = 1;  // the CTA that signs the results must be the only one
    validity_kernel<ROLLOUT_WARPS><<<grid, ROLLOUT_WARPS * 32, smem, stream>>>(p);
    return cudaGetLastError();
}
#endif  // MPS_CPU_EMU
