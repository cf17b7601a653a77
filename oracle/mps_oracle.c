/*
 * mps_oracle.c — CPU oracle (plain C, scalar, sequential) of the kinodynamic-RRT
 * extend step of JoshuaHaustein/manipulation_planning_suite.
 *
 * TEST INFRASTRUCTURE — see mps_oracle.h.  Never linked into the product.
 *
 * Build: gcc -O3 -ffp-contract=off -mfma (no implicit FMA contraction: every +,-,*,/,
 * sqrtf and every explicit fmaf below is one correctly rounded IEEE-754 operation,
 * which is what makes a bit-level comparison against the CUDA path meaningful; the
 * spec says where a fused multiply-add is used).
 *
 * Reference files restated here (relative to the reference root):
 *   src/mps/planner/util/Math.cpp:8-44                        normalize_orientation, shortest_direction_so2
 *   src/mps/planner/ompl/control/RampVelocityControl.cpp:27-158   ramp profile
 *   src/mps/planner/ompl/control/SimEnvStatePropagator.cpp:42-118 propagate loop
 *   src/mps/planner/ompl/control/NaiveControlSampler.cpp:34-71    best-of-K
 *   src/mps/planner/ompl/state/SimEnvWorldStateDistanceMeasure.cpp:34-51  distance
 *   src/mps/planner/ompl/state/SimEnvState.cpp:105-139,1365-1394,1441-1502 state writes, policy, validity
 *   src/mps/planner/ompl/state/goal/ObjectsRelocationGoal.cpp:108-123     goal distance
 *   src/mps/planner/pushing/algorithm/RRT.cpp:449-489,531-565            chains (HybridActionRRT)
 * The physics step (sim_env/Box2D in the reference; absent here) follows the
 * spec "mps-planar-v3" of DESIGN.md §3 — PARITY UNPINNED for that part.
 */
#include "mps_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

#define ORC_PI_D 3.14159265358979323846
#define ORC_PI_F 3.14159265358979323846f
#define ORC_TWO_PI_F (2.0f * 3.14159265358979323846f)

/* ------------------------------------------------------------------------- */
/* util/Math.cpp                                                              */
/* ------------------------------------------------------------------------- */

/* Math.cpp:8-16 (double overload; from ompl SO2StateSpace::enforceBounds) */
void orc_normalize_orientation_d(double* value)
{
    double v = fmod(*value, 2.0 * ORC_PI_D);
    if (v < -ORC_PI_D)
        v += 2.0 * ORC_PI_D;
    else if (v >= ORC_PI_D)
        v -= 2.0 * ORC_PI_D;
    *value = v;
}

/* Math.cpp:18-25 (float overload) */
void orc_normalize_orientation_f(float* value)
{
    float v = fmodf(*value, ORC_TWO_PI_F);
    if (v < -ORC_PI_F)
        v += ORC_TWO_PI_F;
    else if (v >= ORC_PI_F)
        v -= ORC_TWO_PI_F;
    *value = v;
}

/* Math.cpp:34-44 */
float orc_shortest_direction_so2(float val_1, float val_2)
{
    float value = val_2 - val_1;
    if (fabsf(value) > ORC_PI_F) {
        if (value > 0.0f)
            value -= ORC_TWO_PI_F;
        else
            value += ORC_TWO_PI_F;
    }
    return value;
}

/* ------------------------------------------------------------------------- */
/* RampVelocityControl                                                        */
/* ------------------------------------------------------------------------- */

/* setParameters + computeRamp, RampVelocityControl.cpp:34-45,137-158 */
void orc_ramp_set(orc_ramp* r, const float params[5], const float accel_limits[3])
{
    int i;
    float max_accel_time;
    for (i = 0; i < 3; ++i) r->v[i] = params[i];
    r->t_plateau = params[3];
    r->t_rest = params[4];
    max_accel_time = fabsf(r->v[0] / accel_limits[0]);
    for (i = 0; i < 3; ++i) {
        float t = fabsf(r->v[i] / accel_limits[i]);
        max_accel_time = t > max_accel_time ? t : max_accel_time; /* std::max(a, b): b if a < b */
    }
    if (max_accel_time != 0.0f) {
        float inv = 1.0f / max_accel_time;
        for (i = 0; i < 3; ++i) r->a[i] = inv * r->v[i];
    } else {
        for (i = 0; i < 3; ++i) r->a[i] = 0.0f;
    }
    r->t_acc = max_accel_time;
}

/* getVelocity, RampVelocityControl.cpp:98-110.  Note the third comparison is done in
 * double there (the literal 2.0). */
void orc_ramp_velocity(const orc_ramp* r, float t, float vel[3])
{
    int i;
    if (t < r->t_acc) {
        for (i = 0; i < 3; ++i) vel[i] = t * r->a[i];
    } else if (t < r->t_acc + r->t_plateau) {
        for (i = 0; i < 3; ++i) vel[i] = r->v[i];
    } else if ((double)t < 2.0 * (double)r->t_acc + (double)r->t_plateau) {
        float s = t - r->t_acc - r->t_plateau;
        for (i = 0; i < 3; ++i) vel[i] = r->v[i] - s * r->a[i];
    } else {
        for (i = 0; i < 3; ++i) vel[i] = 0.0f;
    }
}

/* getMaxDuration, RampVelocityControl.cpp:112-115 */
float orc_ramp_max_duration(const orc_ramp* r)
{
    return 2.0f * r->t_acc + r->t_plateau + r->t_rest;
}

/* setFromLocalTwist, RampVelocityControl.cpp:68-79 (libm cos/sin: host-side control generation only) */
void orc_ramp_from_local_twist(const float robot_pose[3], const float ltwist[4], float params[5])
{
    float heading = robot_pose[2] + ltwist[3];
    params[0] = ltwist[0] * cosf(heading);
    params[1] = ltwist[0] * sinf(heading);
    params[2] = ltwist[1];
    params[3] = ltwist[2];
    params[4] = 0.0f;
}

/* ------------------------------------------------------------------------- */
/* distance / goal / policy                                                   */
/* ------------------------------------------------------------------------- */

/* SimEnvWorldStateDistanceMeasure.cpp:41-46 with the sub-space distances of
 * OMPL's CompoundStateSpace (sum of weight * component distance), RealVectorStateSpace
 * (sqrt of the sum of squares) and SO2StateSpace (arc) in double, over float-valued
 * states (SimEnvState.cpp:105-135 stores (double)float). */
double orc_distance(const mps_scene* sc, const float* a, const float* b, const float* weights, uint32_t mask)
{
    double dist = 0.0;
    int i;
    for (i = 0; i < sc->n_bodies; ++i) {
        if ((mask >> i) & 1u) {
            double dx = (double)a[3 * i] - (double)b[3 * i];
            double dy = (double)a[3 * i + 1] - (double)b[3 * i + 1];
            double dpos = sqrt(dx * dx + dy * dy);
            double d = fabs((double)a[3 * i + 2] - (double)b[3 * i + 2]);
            double arc = (d > ORC_PI_D) ? 2.0 * ORC_PI_D - d : d;
            double cfg = sc->position_weight * dpos + sc->orientation_weight * arc;
            double w = weights ? (double)weights[i] : 1.0;
            dist += w * cfg;
        }
    }
    return dist;
}

/* ObjectsRelocationGoal.cpp:108-123: the coordinate differences are float subtractions
 * (Eigen float minus Eigen float), squared and summed in double. */
double orc_goal_distance(const float* state, const mps_goal* goals, int n_goals)
{
    double distance = 0.0;
    int g;
    for (g = 0; g < n_goals; ++g) {
        float fx = goals[g].x - state[3 * goals[g].body];
        float fy = goals[g].y - state[3 * goals[g].body + 1];
        double dist_to_center = sqrt((double)fx * (double)fx + (double)fy * (double)fy);
        double e = dist_to_center - (double)goals[g].tol;
        distance += e > 0.0 ? e : 0.0;
    }
    return distance;
}

/* ompl::base::GoalRegion::isSatisfied: distanceGoal(st) <= threshold_, default threshold epsilon */
int orc_goal_satisfied(const float* state, const mps_goal* goals, int n_goals)
{
    if (n_goals <= 0) return 0;
    return orc_goal_distance(state, goals, n_goals) <= DBL_EPSILON;
}

/* CollisionPolicy::collisionAllowed(obj1, obj2), SimEnvState.cpp:1375-1394 */
int orc_collision_allowed(const mps_scene* sc, int a, int b)
{
    int B = sc->n_bodies;
    if (a >= B && b >= B) return 1;
    if (a >= B) return !((sc->static_forbidden >> b) & 1u);
    if (b >= B) return !((sc->static_forbidden >> a) & 1u);
    return !((sc->pair_forbidden[a] >> b) & 1u);
}

/* ------------------------------------------------------------------------- */
/* physics spec mps-planar-v3                                                 */
/* ------------------------------------------------------------------------- */

/* deterministic sincos: Cody-Waite reduction by pi/2, degree-7/6 polynomials */
void orc_sincos(float a, float* s_out, float* c_out)
{
    const float two_over_pi = 0.63661977236758134f;
    const float p_hi = 1.5703125f;
    const float p_mid = 4.837512969970703125e-4f;
    const float p_lo = 7.54978995489188216e-8f;
    float kf = rintf(a * two_over_pi);
    int k = (int)kf;
    float r = fmaf(-kf, p_hi, a);
    float z, sp, cp, sn, cs;
    r = fmaf(-kf, p_mid, r);
    r = fmaf(-kf, p_lo, r);
    z = r * r;
    sp = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    sp = sp * z;
    sp = fmaf(sp, r, r);
    cp = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    cp = cp * z;
    cp = fmaf(cp, z, fmaf(-0.5f, z, 1.0f));
    switch (k & 3) {
    case 0: sn = sp; cs = cp; break;
    case 1: sn = cp; cs = -sp; break;
    case 2: sn = -sp; cs = -cp; break;
    default: sn = -cp; cs = sp; break;
    }
    *s_out = sn;
    *c_out = cs;
}

typedef struct orc_shape {
    int shape;
    float x, y, c, s, hx, hy;
} orc_shape;

typedef struct orc_manifold {
    int a, b; /* b >= B: static b - B */
    int count;
    float nx, ny; /* from a to b */
    float rax[2], ray[2], rbx[2], rby[2], sep[2];
    /* solver rows (spec §3.4): Jacobian lever arms, effective masses, bias, accumulated impulses */
    float rna[2], rnb[2], rta[2], rtb[2];
    float nmass[2], tmass[2], kb[2], acc_n[2], acc_t[2]; /* kb = nmass * bias */
    float ktt, knn, ktn[2][2]; /* couplings between the rows: K = J M^-1 J^T off the diagonal */
    float ma, ia, mb, ib;
    float friction;
} orc_manifold;

typedef struct orc_world {
    const mps_scene* sc;
    int B, S, robot;
    float x[MPS_MAX_BODIES], y[MPS_MAX_BODIES], th[MPS_MAX_BODIES];
    float c[MPS_MAX_BODIES], s[MPS_MAX_BODIES];
    float vx[MPS_MAX_BODIES], vy[MPS_MAX_BODIES], w[MPS_MAX_BODIES];
    float inv_m[MPS_MAX_BODIES], inv_i[MPS_MAX_BODIES], brad[MPS_MAX_BODIES];
    float lin_cap[MPS_MAX_BODIES], ang_cap[MPS_MAX_BODIES];
    float st_c[MPS_MAX_STATICS], st_s[MPS_MAX_STATICS], st_brad[MPS_MAX_STATICS];
    float beta_dt, ref_tol;
    orc_manifold m[MPS_MAX_MANIFOLDS];
    int nm;
    int slot[MPS_MAX_BODIES][MPS_MAX_BODIES + MPS_MAX_STATICS];
    int overflow, bad_contact;
    float max_pen;
    long n_candidates, n_touching, n_colour_passes;
} orc_world;

static float bounding_radius(int shape, float hx, float hy)
{
    if (shape == MPS_SHAPE_DISC) return hx;
    return sqrtf(hx * hx + hy * hy);
}

static void world_init(orc_world* wd, const mps_scene* sc)
{
    int i;
    memset(wd, 0, sizeof(*wd));
    wd->sc = sc;
    wd->B = sc->n_bodies;
    wd->S = sc->n_statics;
    wd->robot = sc->robot_id;
    for (i = 0; i < wd->B; ++i) {
        if (i == wd->robot) {
            wd->inv_m[i] = 0.0f;
            wd->inv_i[i] = 0.0f;
            wd->lin_cap[i] = 0.0f;
            wd->ang_cap[i] = 0.0f;
        } else {
            float mg = sc->mass[i] * sc->gravity;
            float f = sc->mu_ground[i] * mg;
            wd->inv_m[i] = 1.0f / sc->mass[i];
            wd->inv_i[i] = 1.0f / sc->inertia[i];
            wd->lin_cap[i] = sc->dt * f;
            wd->ang_cap[i] = sc->dt * (f * sc->fr_radius[i]);
        }
        wd->brad[i] = bounding_radius(sc->shape[i], sc->hx[i], sc->hy[i]);
    }
    for (i = 0; i < wd->S; ++i) {
        orc_sincos(sc->s_theta[i], &wd->st_s[i], &wd->st_c[i]);
        wd->st_brad[i] = bounding_radius(sc->s_shape[i], sc->s_hx[i], sc->s_hy[i]);
    }
    wd->beta_dt = sc->baumgarte * (1.0f / sc->dt);
    wd->ref_tol = 0.1f * sc->slop;
}

static float wrap_angle(float th)
{
    if (th >= ORC_PI_F)
        th -= ORC_TWO_PI_F;
    else if (th < -ORC_PI_F)
        th += ORC_TWO_PI_F;
    return th;
}

static void world_set_state(orc_world* wd, const float* state)
{
    int i;
    for (i = 0; i < wd->B; ++i) {
        wd->x[i] = state[3 * i];
        wd->y[i] = state[3 * i + 1];
        wd->th[i] = wrap_angle(state[3 * i + 2]);
        orc_sincos(wd->th[i], &wd->s[i], &wd->c[i]);
        wd->vx[i] = 0.0f;
        wd->vy[i] = 0.0f;
        wd->w[i] = 0.0f;
    }
}

static orc_shape body_shape(const orc_world* wd, int i)
{
    orc_shape sh;
    sh.shape = wd->sc->shape[i];
    sh.x = wd->x[i];
    sh.y = wd->y[i];
    sh.c = wd->c[i];
    sh.s = wd->s[i];
    sh.hx = wd->sc->hx[i];
    sh.hy = wd->sc->hy[i];
    return sh;
}

static orc_shape static_shape(const orc_world* wd, int k)
{
    orc_shape sh;
    sh.shape = wd->sc->s_shape[k];
    sh.x = wd->sc->s_x[k];
    sh.y = wd->sc->s_y[k];
    sh.c = wd->st_c[k];
    sh.s = wd->st_s[k];
    sh.hx = wd->sc->s_hx[k];
    sh.hy = wd->sc->s_hy[k];
    return sh;
}

/* ---- narrow phase ------------------------------------------------------- */

static int collide_disc_disc(const orc_shape* A, const orc_shape* Bs, orc_manifold* m)
{
    float dx = Bs->x - A->x;
    float dy = Bs->y - A->y;
    float d2 = fmaf(dx, dx, dy * dy);
    float R = A->hx + Bs->hx;
    float d, nx, ny, sep, h;
    if (d2 > R * R) return 0;
    d = sqrtf(d2);
    if (d > 1e-9f) {
        nx = dx / d;
        ny = dy / d;
    } else {
        nx = 1.0f;
        ny = 0.0f;
    }
    sep = d - R;
    h = fmaf(0.5f, sep, A->hx);
    m->nx = nx;
    m->ny = ny;
    m->count = 1;
    m->rax[0] = nx * h;
    m->ray[0] = ny * h;
    m->rbx[0] = m->rax[0] - dx;
    m->rby[0] = m->ray[0] - dy;
    m->sep[0] = sep;
    return 1;
}

/* box X against disc D.  a_is_box: body a of the pair is the box. */
static int collide_box_disc(const orc_shape* X, const orc_shape* D, int a_is_box, orc_manifold* m)
{
    float ddx = D->x - X->x;
    float ddy = D->y - X->y;
    float lx = fmaf(ddx, X->c, ddy * X->s);
    float ly = fmaf(ddy, X->c, -(ddx * X->s));
    float qx = fminf(fmaxf(lx, -X->hx), X->hx);
    float qy = fminf(fmaxf(ly, -X->hy), X->hy);
    float r = D->hx;
    float nlx, nly, sep, spx, spy, cplx, cply, rxx, rxy, nx, ny, rdx, rdy, hs;
    if (qx == lx && qy == ly) {
        float px = X->hx - fabsf(lx);
        float py = X->hy - fabsf(ly);
        if (px < py) {
            nlx = lx >= 0.0f ? 1.0f : -1.0f;
            nly = 0.0f;
            sep = -(px + r);
            spx = nlx * X->hx;
            spy = ly;
        } else {
            nlx = 0.0f;
            nly = ly >= 0.0f ? 1.0f : -1.0f;
            sep = -(py + r);
            spx = lx;
            spy = nly * X->hy;
        }
    } else {
        float vx = lx - qx;
        float vy = ly - qy;
        float d2 = fmaf(vx, vx, vy * vy);
        float d;
        if (d2 > r * r) return 0;
        d = sqrtf(d2);
        nlx = vx / d;
        nly = vy / d;
        sep = d - r;
        spx = qx;
        spy = qy;
    }
    hs = 0.5f * sep;
    cplx = fmaf(nlx, hs, spx);
    cply = fmaf(nly, hs, spy);
    rxx = fmaf(X->c, cplx, -(X->s * cply));
    rxy = fmaf(X->s, cplx, X->c * cply);
    nx = fmaf(X->c, nlx, -(X->s * nly));
    ny = fmaf(X->s, nlx, X->c * nly);
    rdx = rxx - ddx;
    rdy = rxy - ddy;
    m->count = 1;
    m->sep[0] = sep;
    if (a_is_box) {
        m->nx = nx;
        m->ny = ny;
        m->rax[0] = rxx;
        m->ray[0] = rxy;
        m->rbx[0] = rdx;
        m->rby[0] = rdy;
    } else {
        m->nx = -nx;
        m->ny = -ny;
        m->rax[0] = rdx;
        m->ray[0] = rdy;
        m->rbx[0] = rxx;
        m->rby[0] = rxy;
    }
    return 1;
}

static int collide_box_box(const orc_shape* A, const orc_shape* Bs, float ref_tol, orc_manifold* m)
{
    float dx = Bs->x - A->x;
    float dy = Bs->y - A->y;
    float d00 = fmaf(A->c, Bs->c, A->s * Bs->s);
    float d01 = fmaf(A->s, Bs->c, -(A->c * Bs->s));
    float ad00 = fabsf(d00);
    float ad01 = fabsf(d01);
    float pax = fmaf(dx, A->c, dy * A->s);
    float pay = fmaf(dy, A->c, -(dx * A->s));
    float pbx = fmaf(dx, Bs->c, dy * Bs->s);
    float pby = fmaf(dy, Bs->c, -(dx * Bs->s));
    float sep_ax = fabsf(pax) - A->hx - fmaf(Bs->hx, ad00, Bs->hy * ad01);
    float sep_ay = fabsf(pay) - A->hy - fmaf(Bs->hx, ad01, Bs->hy * ad00);
    float sep_bx = fabsf(pbx) - Bs->hx - fmaf(A->hx, ad00, A->hy * ad01);
    float sep_by = fabsf(pby) - Bs->hy - fmaf(A->hx, ad01, A->hy * ad00);
    int a_face_y, b_face_y, ref_is_b, face_y, inc_y, n, k;
    float sep_a, sep_b;
    const orc_shape *R, *I;
    float ddx, ddy, ux, uy, hu, ht, proj, sgn, nrx, nry, trx, tryy, e0, e1, sm, mx, my, hm, he;
    float cx, cy, ex, ey, v1x, v1y, v2x, v2y, t1, t2, s1, s2, ct[2], cs[2];
    if (sep_ax > 0.0f || sep_ay > 0.0f || sep_bx > 0.0f || sep_by > 0.0f) return 0;
    a_face_y = sep_ay > sep_ax;
    b_face_y = sep_by > sep_bx;
    sep_a = a_face_y ? sep_ay : sep_ax;
    sep_b = b_face_y ? sep_by : sep_bx;
    ref_is_b = sep_b > sep_a + ref_tol;
    R = ref_is_b ? Bs : A;
    I = ref_is_b ? A : Bs;
    ddx = ref_is_b ? -dx : dx;
    ddy = ref_is_b ? -dy : dy;
    face_y = ref_is_b ? b_face_y : a_face_y;
    ux = face_y ? -R->s : R->c;
    uy = face_y ? R->c : R->s;
    hu = face_y ? R->hy : R->hx;
    ht = face_y ? R->hx : R->hy;
    proj = fmaf(ddx, ux, ddy * uy);
    sgn = proj >= 0.0f ? 1.0f : -1.0f;
    nrx = sgn * ux;
    nry = sgn * uy;
    trx = -nry;
    tryy = nrx;
    /* incident face of I: the face normal most opposed to the reference normal */
    e0 = fmaf(nrx, I->c, nry * I->s);
    e1 = fmaf(nry, I->c, -(nrx * I->s));
    inc_y = !(fabsf(e0) >= fabsf(e1));
    if (!inc_y) {
        sm = e0 > 0.0f ? -1.0f : 1.0f;
        mx = sm * I->c;
        my = sm * I->s;
        hm = I->hx;
        he = I->hy;
    } else {
        sm = e1 > 0.0f ? -1.0f : 1.0f;
        mx = sm * -I->s;
        my = sm * I->c;
        hm = I->hy;
        he = I->hx;
    }
    cx = fmaf(mx, hm, ddx);
    cy = fmaf(my, hm, ddy);
    ex = -my;
    ey = mx;
    v1x = fmaf(ex, he, cx);
    v1y = fmaf(ey, he, cy);
    v2x = fmaf(-ex, he, cx);
    v2y = fmaf(-ey, he, cy);
    t1 = fmaf(v1x, trx, v1y * tryy);
    t2 = fmaf(v2x, trx, v2y * tryy);
    s1 = fmaf(v1x, nrx, v1y * nry) - hu;
    s2 = fmaf(v2x, nrx, v2y * nry) - hu;
    if ((t1 > ht && t2 > ht) || (t1 < -ht && t2 < -ht)) return 0;
    /* clip both end points of the incident edge to the side planes |t| <= ht */
    ct[0] = t1;
    cs[0] = s1;
    ct[1] = t2;
    cs[1] = s2;
    if (t1 > ht) {
        float al = (ht - t1) / (t2 - t1);
        cs[0] = fmaf(al, s2 - s1, s1);
        ct[0] = ht;
    } else if (t1 < -ht) {
        float al = (-ht - t1) / (t2 - t1);
        cs[0] = fmaf(al, s2 - s1, s1);
        ct[0] = -ht;
    }
    if (t2 > ht) {
        float al = (ht - t2) / (t1 - t2);
        cs[1] = fmaf(al, s1 - s2, s2);
        ct[1] = ht;
    } else if (t2 < -ht) {
        float al = (-ht - t2) / (t1 - t2);
        cs[1] = fmaf(al, s1 - s2, s2);
        ct[1] = -ht;
    }
    n = 0;
    for (k = 0; k < 2; ++k) {
        if (cs[k] <= 0.0f) {
            float h = fmaf(0.5f, cs[k], hu);
            float rrx = fmaf(trx, ct[k], nrx * h);
            float rry = fmaf(tryy, ct[k], nry * h);
            float rix = rrx - ddx;
            float riy = rry - ddy;
            if (ref_is_b) {
                m->rax[n] = rix;
                m->ray[n] = riy;
                m->rbx[n] = rrx;
                m->rby[n] = rry;
            } else {
                m->rax[n] = rrx;
                m->ray[n] = rry;
                m->rbx[n] = rix;
                m->rby[n] = riy;
            }
            m->sep[n] = cs[k];
            ++n;
        }
    }
    if (n == 0) return 0;
    m->count = n;
    m->nx = ref_is_b ? -nrx : nrx;
    m->ny = ref_is_b ? -nry : nry;
    return n;
}

static int collide(const orc_shape* A, const orc_shape* Bs, float ref_tol, orc_manifold* m)
{
    if (A->shape == MPS_SHAPE_BOX && Bs->shape == MPS_SHAPE_BOX) return collide_box_box(A, Bs, ref_tol, m);
    if (A->shape == MPS_SHAPE_BOX && Bs->shape == MPS_SHAPE_DISC) return collide_box_disc(A, Bs, 1, m);
    if (A->shape == MPS_SHAPE_DISC && Bs->shape == MPS_SHAPE_BOX) return collide_box_disc(Bs, A, 0, m);
    return collide_disc_disc(A, Bs, m);
}

/* solver constants of one manifold: Jacobian rows, effective masses, bias, mixed friction */
static void manifold_prepare(const orc_world* wd, orc_manifold* m)
{
    const mps_scene* sc = wd->sc;
    int B = wd->B, k;
    float ma = wd->inv_m[m->a], ia = wd->inv_i[m->a];
    float mb = 0.0f, ib = 0.0f, mub;
    float tx = m->ny, ty = -m->nx;
    if (m->b < B) {
        mb = wd->inv_m[m->b];
        ib = wd->inv_i[m->b];
        mub = sc->mu_contact[m->b];
    } else {
        mub = sc->s_mu_contact[m->b - B];
    }
    m->friction = sqrtf(sc->mu_contact[m->a] * mub);
    m->ma = ma;
    m->ia = ia;
    m->mb = mb;
    m->ib = ib;
    {
        float iarna[2] = {0.0f, 0.0f}, ibrnb[2] = {0.0f, 0.0f}, iarta[2] = {0.0f, 0.0f}, ibrtb[2] = {0.0f, 0.0f};
        for (k = 0; k < m->count; ++k) {
            float rna = fmaf(m->rax[k], m->ny, -(m->ray[k] * m->nx));
            float rnb = fmaf(m->rbx[k], m->ny, -(m->rby[k] * m->nx));
            float rta = fmaf(m->rax[k], ty, -(m->ray[k] * tx));
            float rtb = fmaf(m->rbx[k], ty, -(m->rby[k] * tx));
            float kn, kt, pen, nmass, bias;
            iarna[k] = ia * rna;
            ibrnb[k] = ib * rnb;
            iarta[k] = ia * rta;
            ibrtb[k] = ib * rtb;
            kn = fmaf(ibrnb[k], rnb, fmaf(iarna[k], rna, ma + mb));
            kt = fmaf(ibrtb[k], rtb, fmaf(iarta[k], rta, ma + mb));
            pen = -m->sep[k] - sc->slop;
            m->rna[k] = rna;
            m->rnb[k] = rnb;
            m->rta[k] = rta;
            m->rtb[k] = rtb;
            nmass = kn > 0.0f ? 1.0f / kn : 0.0f;
            bias = pen > 0.0f ? fminf(wd->beta_dt * pen, sc->max_bias_vel) : 0.0f;
            m->nmass[k] = nmass;
            m->tmass[k] = kt > 0.0f ? 1.0f / kt : 0.0f;
            m->kb[k] = nmass * bias;
            m->acc_n[k] = 0.0f;
            m->acc_t[k] = 0.0f;
        }
        /* how an impulse on one row changes the relative velocity along another (spec §3.4) */
        m->ktn[0][0] = fmaf(iarta[0], m->rna[0], ibrtb[0] * m->rnb[0]);
        m->ktt = 0.0f;
        m->knn = 0.0f;
        m->ktn[0][1] = m->ktn[1][0] = m->ktn[1][1] = 0.0f;
        if (m->count > 1) {
            m->ktt = fmaf(ibrtb[0], m->rtb[1], fmaf(iarta[0], m->rta[1], ma + mb));
            m->knn = fmaf(ibrnb[0], m->rnb[1], fmaf(iarna[0], m->rna[1], ma + mb));
            m->ktn[0][1] = fmaf(iarta[0], m->rna[1], ibrtb[0] * m->rnb[1]);
            m->ktn[1][0] = fmaf(iarta[1], m->rna[0], ibrtb[1] * m->rnb[0]);
            m->ktn[1][1] = fmaf(iarta[1], m->rna[1], ibrtb[1] * m->rnb[1]);
        }
    }
}

/* collision detection at the current poses.  Fills wd->m / wd->slot, sets overflow,
 * bad_contact (a touching pair the policy forbids) and max_pen. */
static void world_detect(orc_world* wd)
{
    int B = wd->B, S = wd->S, i, j;
    wd->nm = 0;
    wd->bad_contact = 0;
    wd->max_pen = 0.0f;
    for (i = 0; i < B; ++i)
        for (j = 0; j < B + S; ++j) wd->slot[i][j] = -1;
    for (i = 0; i < B; ++i) {
        orc_shape A = body_shape(wd, i);
        for (j = i + 1; j < B + S; ++j) {
            orc_shape Bs;
            float rr, dx, dy;
            orc_manifold m;
            int k;
            /* broad phase: conservative filters only (they never change the result) */
            if (j < B) {
                Bs = body_shape(wd, j);
                rr = wd->brad[i] + wd->brad[j];
                dx = Bs.x - A.x;
                dy = Bs.y - A.y;
                if (dx * dx + dy * dy > rr * rr) continue;
            } else {
                Bs = static_shape(wd, j - B);
                dx = A.x - Bs.x;
                dy = A.y - Bs.y;
                rr = wd->brad[i] * 1.001f + 1e-5f;
                if (Bs.shape == MPS_SHAPE_BOX) {
                    float lx = dx * Bs.c + dy * Bs.s;
                    float ly = dy * Bs.c - dx * Bs.s;
                    float ex = fmaxf(fabsf(lx) - Bs.hx, 0.0f);
                    float ey = fmaxf(fabsf(ly) - Bs.hy, 0.0f);
                    if (ex * ex + ey * ey > rr * rr) continue;
                } else {
                    rr += Bs.hx;
                    if (dx * dx + dy * dy > rr * rr) continue;
                }
            }
            wd->n_candidates++;
            memset(&m, 0, sizeof(m));
            m.a = i;
            m.b = j;
            if (!collide(&A, &Bs, wd->ref_tol, &m)) continue;
            wd->n_touching++;
            for (k = 0; k < m.count; ++k)
                if (-m.sep[k] > wd->max_pen) wd->max_pen = -m.sep[k];
            if (!orc_collision_allowed(wd->sc, i, j)) wd->bad_contact = 1;
            if (wd->nm >= wd->sc->max_manifolds) {
                wd->overflow = 1;
                continue;
            }
            manifold_prepare(wd, &m);
            wd->slot[i][j] = wd->nm;
            wd->m[wd->nm++] = m;
        }
    }
}

/* One Gauss-Seidel visit of a manifold: friction rows (point 0, point 1), then the non-penetration rows.
 * The relative velocities along all rows are evaluated once from the twists at the start of the visit
 * (fused multiply-add chains, one rounding per step); an impulse d on an earlier row reaches a later row
 * through the coupling coefficient, u_s += K_sr d_r, which is what re-evaluating u_s from the updated
 * twists would give in exact arithmetic.  Impulse increments are clamped directly
 * (d = clamp(-u / K_ss, lo - acc, hi - acc)) and the twists are updated once at the end (spec §3.4). */
static void solve_manifold(orc_world* wd, orc_manifold* m)
{
    int B = wd->B, k;
    int a = m->a, b = m->b;
    float vax = wd->vx[a], vay = wd->vy[a], wa = wd->w[a];
    float vbx = 0.0f, vby = 0.0f, wb = 0.0f;
    float nx = m->nx, ny = m->ny, tx = m->ny, ty = -m->nx;
    float ut[2], un[2], dt_[2], dn[2];
    float dvx, dvy, px, py, ra, rb, st;
    if (b < B) {
        vbx = wd->vx[b];
        vby = wd->vy[b];
        wb = wd->w[b];
    }
    dvx = vbx - vax;
    dvy = vby - vay;
    /* A manifold with one point runs the same arithmetic with the second point's coefficients all zero
     * (manifold_prepare leaves them so): no data-dependent branch on the device, identical bits here. */
    for (k = 0; k < 2; ++k) {
        ut[k] = fmaf(dvx, tx, fmaf(dvy, ty, fmaf(wb, m->rtb[k], -(wa * m->rta[k]))));
        un[k] = fmaf(dvx, nx, fmaf(dvy, ny, fmaf(wb, m->rnb[k], -(wa * m->rna[k]))));
    }
    { /* friction, point 0 */
        float lim = m->friction * m->acc_n[0];
        float dl = m->tmass[0] * -ut[0];
        float d = fminf(fmaxf(dl, -lim - m->acc_t[0]), lim - m->acc_t[0]);
        dt_[0] = d;
        m->acc_t[0] = m->acc_t[0] + d;
    }
    { /* friction, point 1 */
        float u = fmaf(m->ktt, dt_[0], ut[1]);
        float lim = m->friction * m->acc_n[1];
        float dl = m->tmass[1] * -u;
        float d = fminf(fmaxf(dl, -lim - m->acc_t[1]), lim - m->acc_t[1]);
        dt_[1] = d;
        m->acc_t[1] = m->acc_t[1] + d;
    }
    { /* non-penetration, point 0 */
        float u = fmaf(m->ktn[1][0], dt_[1], fmaf(m->ktn[0][0], dt_[0], un[0]));
        float dl = fmaf(-m->nmass[0], u, m->kb[0]);
        float d = fmaxf(dl, -m->acc_n[0]);
        dn[0] = d;
        m->acc_n[0] = m->acc_n[0] + d;
    }
    { /* non-penetration, point 1 */
        float u = fmaf(m->knn, dn[0], fmaf(m->ktn[1][1], dt_[1], fmaf(m->ktn[0][1], dt_[0], un[1])));
        float dl = fmaf(-m->nmass[1], u, m->kb[1]);
        float d = fmaxf(dl, -m->acc_n[1]);
        dn[1] = d;
        m->acc_n[1] = m->acc_n[1] + d;
    }
    /* total impulse on b (minus that on a) and the angular parts */
    st = dt_[0] + dt_[1];
    ra = fmaf(m->rna[1], dn[1], fmaf(m->rta[1], dt_[1], fmaf(m->rna[0], dn[0], m->rta[0] * dt_[0])));
    rb = fmaf(m->rnb[1], dn[1], fmaf(m->rtb[1], dt_[1], fmaf(m->rnb[0], dn[0], m->rtb[0] * dt_[0])));
    px = fmaf(nx, dn[1], fmaf(nx, dn[0], tx * st));
    py = fmaf(ny, dn[1], fmaf(ny, dn[0], ty * st));
    wd->vx[a] = fmaf(-m->ma, px, vax);
    wd->vy[a] = fmaf(-m->ma, py, vay);
    wd->w[a] = fmaf(-m->ia, ra, wa);
    if (b < B) {
        wd->vx[b] = fmaf(m->mb, px, vbx);
        wd->vy[b] = fmaf(m->mb, py, vby);
        wd->w[b] = fmaf(m->ib, rb, wb);
    }
}

/* one physics step with robot target velocity u.  Returns 1 when the step saw a contact
 * the collision policy forbids (the reference's stepPhysics(callback) == false). */
static int world_step(orc_world* wd, const float u[3])
{
    const mps_scene* sc = wd->sc;
    int B = wd->B, S = wd->S, i, it, col;
    wd->vx[wd->robot] = u[0];
    wd->vy[wd->robot] = u[1];
    wd->w[wd->robot] = u[2];
    world_detect(wd);
    if (wd->overflow) return wd->bad_contact;
    /* support-plane friction, once per step before the contact iterations: the impulse that brings
     * the body to rest, capped at dt * mu * m * g (linear, vector norm) and dt * mu * m * g * r (angular) */
    for (i = 0; i < B; ++i) {
        float imp, impc, ix, iy, n2, cap;
        if (i == wd->robot) continue;
        imp = -sc->inertia[i] * wd->w[i];
        impc = fminf(fmaxf(imp, -wd->ang_cap[i]), wd->ang_cap[i]);
        /* an impulse within the cap stops the body exactly */
        wd->w[i] = impc == imp ? 0.0f : fmaf(wd->inv_i[i], impc, wd->w[i]);
        ix = -sc->mass[i] * wd->vx[i];
        iy = -sc->mass[i] * wd->vy[i];
        n2 = fmaf(ix, ix, iy * iy);
        cap = wd->lin_cap[i];
        if (n2 > cap * cap) {
            float sc_ = cap / sqrtf(n2);
            wd->vx[i] = fmaf(wd->inv_m[i], ix * sc_, wd->vx[i]);
            wd->vy[i] = fmaf(wd->inv_m[i], iy * sc_, wd->vy[i]);
        } else {
            wd->vx[i] = 0.0f;
            wd->vy[i] = 0.0f;
        }
    }
    for (it = 0; it < sc->vel_iters && wd->nm > 0; ++it) {
        /* contacts in colour order: colour r < B holds the body pairs with (i + j) mod B == r,
         * colour B + s the pairs (i, static s); inside a colour, ascending i */
        for (col = 0; col < B + S; ++col) {
            int any = 0;
            for (i = 0; i < B; ++i) {
                int sl;
                if (col < B) {
                    int j = ((col - i) % B + B) % B;
                    if (j <= i) continue;
                    sl = wd->slot[i][j];
                } else {
                    sl = wd->slot[i][col];
                }
                if (sl < 0) continue;
                solve_manifold(wd, &wd->m[sl]);
                any = 1;
            }
            wd->n_colour_passes += any;
        }
    }
    for (i = 0; i < B; ++i) {
        wd->x[i] = fmaf(sc->dt, wd->vx[i], wd->x[i]);
        wd->y[i] = fmaf(sc->dt, wd->vy[i], wd->y[i]);
        wd->th[i] = wrap_angle(fmaf(sc->dt, wd->w[i], wd->th[i]));
        orc_sincos(wd->th[i], &wd->s[i], &wd->c[i]);
    }
    return wd->bad_contact;
}

static int world_at_rest(const orc_world* wd)
{
    const mps_scene* sc = wd->sc;
    float v2max = sc->v_rest * sc->v_rest;
    int i;
    for (i = 0; i < wd->B; ++i) {
        if (fmaf(wd->vx[i], wd->vx[i], wd->vy[i] * wd->vy[i]) > v2max) return 0;
        if (fabsf(wd->w[i]) > sc->w_rest) return 0;
    }
    return 1;
}

/* extractState (SimEnvState.cpp:1199-1208) -> setConfiguration normalises theta in double
 * (SimEnvState.cpp:131-135), getConfiguration reads it back as float (:96-98) */
static void world_extract(const orc_world* wd, float* state)
{
    int i;
    for (i = 0; i < wd->B; ++i) {
        double th = (double)wd->th[i];
        orc_normalize_orientation_d(&th);
        state[3 * i] = wd->x[i];
        state[3 * i + 1] = wd->y[i];
        state[3 * i + 2] = (float)th;
    }
}

/* satisfiesBounds: RealVectorStateSpace (value - eps > high || value + eps < low -> false),
 * SO2StateSpace (value < pi + eps && value > -pi - eps) */
static int state_in_bounds(const mps_scene* sc, const float* state)
{
    int i;
    for (i = 0; i < sc->n_bodies; ++i) {
        double x = (double)state[3 * i], y = (double)state[3 * i + 1], th = (double)state[3 * i + 2];
        if (x - DBL_EPSILON > (double)sc->x_max || x + DBL_EPSILON < (double)sc->x_min) return 0;
        if (y - DBL_EPSILON > (double)sc->y_max || y + DBL_EPSILON < (double)sc->y_min) return 0;
        if (!(th < ORC_PI_D + DBL_EPSILON && th > -ORC_PI_D - DBL_EPSILON)) return 0;
    }
    return 1;
}

static int world_state_valid(orc_world* wd, const float* state, uint32_t* flags)
{
    uint32_t f = 0;
    if (!state_in_bounds(wd->sc, state)) {
        *flags |= MPS_F_BOUNDS;
        return 0;
    }
    world_set_state(wd, state);
    wd->overflow = 0;
    world_detect(wd);
    if (wd->max_pen > wd->sc->pen_max) f |= MPS_F_INFEASIBLE;
    if (wd->bad_contact) f |= MPS_F_BAD_CONTACT;
    if (wd->overflow) f |= MPS_F_OVERFLOW;
    *flags |= f;
    return f == 0;
}

int orc_is_valid(const mps_scene* sc, const float* state, uint32_t* flags)
{
    orc_world wd;
    uint32_t f = 0;
    int ok;
    world_init(&wd, sc);
    ok = world_state_valid(&wd, state, &f);
    if (flags) *flags = f;
    return ok;
}

static int propagate_impl(const mps_scene* sc, const float* state_in, float* control, float* state_out,
                          uint32_t* flags_out, int32_t* n_steps_out, float* trace, int32_t max_trace,
                          long counters[3])
{
    orc_world wd;
    orc_ramp ramp;
    uint32_t flags = MPS_F_RAN;
    float time = 0.0f, T, u[3];
    int ok = 1, steps = 0, i;
    float params[5];
    world_init(&wd, sc);
    world_set_state(&wd, state_in);
    /* semi-dynamic: rest time reset before propagating (SimEnvStatePropagator.cpp:63-67) */
    for (i = 0; i < 5; ++i) params[i] = control[i];
    params[4] = 0.0f;
    orc_ramp_set(&ramp, params, sc->accel_limits);
    T = orc_ramp_max_duration(&ramp);
    /* active phase (:75-83): the inner `bool propagation_success` shadows the outer one, so
     * contacts never stop this loop unless strict_active_contacts is set */
    while (time < T && ok) {
        int bad;
        orc_ramp_velocity(&ramp, time, u);
        bad = world_step(&wd, u);
        time += sc->dt;
        if (trace && steps < max_trace) {
            for (i = 0; i < wd.B; ++i) {
                float* t = trace + ((long)steps * wd.B + i) * 6;
                t[0] = wd.x[i]; t[1] = wd.y[i]; t[2] = wd.th[i];
                t[3] = wd.vx[i]; t[4] = wd.vy[i]; t[5] = wd.w[i];
            }
        }
        ++steps;
        if (wd.overflow) {
            flags |= MPS_F_OVERFLOW;
            ok = 0;
        }
        if (bad && sc->strict_active_contacts) {
            flags |= MPS_F_ACTIVE_CONTACT;
            ok = 0;
        }
    }
    /* rest phase (:86-105) */
    if (ok) {
        float resting_time = 0.0f;
        int jam = 0;
        u[0] = u[1] = u[2] = 0.0f;
        while (resting_time <= sc->t_max && !world_at_rest(&wd) && ok) {
            int bad = world_step(&wd, u);
            ok = !bad;
            resting_time += sc->dt;
            if (trace && steps < max_trace) {
                for (i = 0; i < wd.B; ++i) {
                    float* t = trace + ((long)steps * wd.B + i) * 6;
                    t[0] = wd.x[i]; t[1] = wd.y[i]; t[2] = wd.th[i];
                    t[3] = wd.vx[i]; t[4] = wd.vy[i]; t[5] = wd.w[i];
                }
            }
            ++steps;
            if (bad) flags |= MPS_F_REST_CONTACT;
            if (wd.overflow) {
                flags |= MPS_F_OVERFLOW;
                ok = 0;
            }
            /* jam rule (spec §3.6): the robot has stopped.  A free contact sheds its penetration beyond
             * the slop by the factor (1 - baumgarte) per step, so even from pen_max it is within 0.25 * slop of the
             * slop after 13 steps: a world that stays deeper than 1.25 * slop for jam_steps (25) consecutive
             * steps is wedged between bodies that do not give way and would creep until t_max */
            jam = wd.max_pen > 1.25f * sc->slop ? jam + 1 : 0;
            if (sc->jam_steps > 0 && jam >= sc->jam_steps) {
                flags |= MPS_F_JAMMED;
                ok = 0;
            }
        }
        control[4] = resting_time;
        if (!(resting_time <= sc->t_max && world_at_rest(&wd))) flags |= MPS_F_NOT_AT_REST;
        ok = resting_time <= sc->t_max && world_at_rest(&wd) && ok;
    } else {
        control[4] = 0.0f;
    }
    world_extract(&wd, state_out);
    {
        uint32_t vf = 0;
        long c0 = wd.n_candidates, c1 = wd.n_touching, c2 = wd.n_colour_passes;
        int valid = world_state_valid(&wd, state_out, &vf); /* :108 */
        wd.n_candidates = c0; wd.n_touching = c1; wd.n_colour_passes = c2;
        flags |= vf;
        ok = ok && valid;
    }
    if (ok) flags |= MPS_F_OK;
    if (flags_out) *flags_out = flags;
    if (n_steps_out) *n_steps_out = steps;
    if (counters) {
        counters[0] += wd.n_candidates;
        counters[1] += wd.n_touching;
        counters[2] += wd.n_colour_passes;
    }
    return ok;
}

int orc_propagate(const mps_scene* sc, const float* state_in, float* control, float* state_out,
                  uint32_t* flags, int32_t* n_steps)
{
    float in_copy[3 * MPS_MAX_BODIES];
    memcpy(in_copy, state_in, sizeof(float) * 3 * (size_t)sc->n_bodies); /* in/out may alias */
    return propagate_impl(sc, in_copy, control, state_out, flags, n_steps, 0, 0, 0);
}

int orc_propagate_trace(const mps_scene* sc, const float* state_in, float* control, float* trace,
                        int32_t max_steps, int32_t* n_steps)
{
    float out[3 * MPS_MAX_BODIES];
    uint32_t flags;
    return propagate_impl(sc, state_in, control, out, &flags, n_steps, trace, max_steps, 0);
}

int orc_step_once(const mps_scene* sc, float* pose, float* vel, const float u[3], int32_t* n_manifolds)
{
    orc_world wd;
    int i, bad;
    world_init(&wd, sc);
    world_set_state(&wd, pose);
    for (i = 0; i < wd.B; ++i) {
        wd.vx[i] = vel[3 * i];
        wd.vy[i] = vel[3 * i + 1];
        wd.w[i] = vel[3 * i + 2];
    }
    bad = world_step(&wd, u);
    for (i = 0; i < wd.B; ++i) {
        pose[3 * i] = wd.x[i];
        pose[3 * i + 1] = wd.y[i];
        pose[3 * i + 2] = wd.th[i];
        vel[3 * i] = wd.vx[i];
        vel[3 * i + 1] = wd.vy[i];
        vel[3 * i + 2] = wd.w[i];
    }
    if (n_manifolds) *n_manifolds = wd.nm;
    return bad;
}

int orc_contacts(const mps_scene* sc, const float* state, int32_t* ab, float* data, int32_t max_m)
{
    orc_world wd;
    int i;
    world_init(&wd, sc);
    world_set_state(&wd, state);
    world_detect(&wd);
    for (i = 0; i < wd.nm && i < max_m; ++i) {
        ab[2 * i] = wd.m[i].a;
        ab[2 * i + 1] = wd.m[i].b;
        data[5 * i] = (float)wd.m[i].count;
        data[5 * i + 1] = wd.m[i].nx;
        data[5 * i + 2] = wd.m[i].ny;
        data[5 * i + 3] = wd.m[i].sep[0];
        data[5 * i + 4] = wd.m[i].count > 1 ? wd.m[i].sep[1] : 0.0f;
    }
    return wd.nm;
}

/* ------------------------------------------------------------------------- */
/* extend: K chains, best-of                                                  */
/* ------------------------------------------------------------------------- */

int orc_extend(const mps_scene* sc, const mps_extend_in* in, const mps_extend_out* out)
{
    int B = sc->n_bodies, K = in->K, L = in->L, k, l, i;
    int best_k = -1, best_nok = 0, best_goal_step = -1;
    double best_d = DBL_MAX;
    float best_f = FLT_MAX;
    float cur[3 * MPS_MAX_BODIES], res[3 * MPS_MAX_BODIES];
    /* scratch for the running chain */
    static __thread float chain_states[MPS_MAX_CHAIN][3 * MPS_MAX_BODIES];
    float chain_ctrl[MPS_MAX_CHAIN][5];
    uint32_t chain_flags[MPS_MAX_CHAIN];
    if (L > MPS_MAX_CHAIN || L < 1 || K < 0) return MPS_ERR_ARG;
    for (k = 0; k < K; ++k) {
        int n = in->n_ctrl ? in->n_ctrl[k] : L;
        int n_ok = 0, goal_step = -1;
        double d = INFINITY;
        if (n > L) n = L;
        memcpy(cur, in->starts ? in->starts + (size_t)k * 3 * B : in->start, sizeof(float) * 3 * (size_t)B);
        for (l = 0; l < L; ++l) {
            chain_flags[l] = 0;
            for (i = 0; i < 5; ++i) chain_ctrl[l][i] = in->controls[((size_t)k * L + l) * 5 + i];
            for (i = 0; i < 3 * B; ++i) chain_states[l][i] = 0.0f;
        }
        /* forwardPropagateActionSequence RRT.cpp:531-565: abort at first failure */
        for (l = 0; l < n; ++l) {
            uint32_t f = 0;
            int32_t steps = 0;
            int ok = propagate_impl(sc, cur, chain_ctrl[l], res, &f, &steps, 0, 0, 0);
            if (ok) {
                if (in->goals && orc_goal_satisfied(res, in->goals, in->n_goals)) {
                    f |= MPS_F_GOAL;
                    if (goal_step < 0) goal_step = l;
                }
                if (in->ball_center) {
                    uint32_t arr_mask = ~(1u << sc->robot_id);
                    if (B < 32) arr_mask &= (1u << B) - 1u;
                    if (orc_distance(sc, res, in->ball_center, in->weights, arr_mask) <= (double)in->ball_radius)
                        f |= MPS_F_IN_BALL;
                }
            }
            chain_flags[l] = f;
            memcpy(chain_states[l], res, sizeof(float) * 3 * (size_t)B);
            if (out->all_steps) out->all_steps[(size_t)k * L + l] = steps;
            if (!ok) break;
            ++n_ok;
            memcpy(cur, res, sizeof(float) * 3 * (size_t)B);
            if (in->stop_at_goal && goal_step >= 0) break; /* forwardPropagatePath RRT.cpp:966-968 */
        }
        if (n_ok > 0 && in->dest) d = orc_distance(sc, chain_states[n_ok - 1], in->dest, in->weights, in->mask);
        for (l = 0; l < L; ++l) {
            if (out->all_states)
                memcpy(out->all_states + ((size_t)k * L + l) * 3 * B, chain_states[l], sizeof(float) * 3 * (size_t)B);
            if (out->all_rest) out->all_rest[(size_t)k * L + l] = chain_ctrl[l][4];
            if (out->all_flags) out->all_flags[(size_t)k * L + l] = chain_flags[l];
            if (out->all_steps && !(chain_flags[l] & MPS_F_RAN)) out->all_steps[(size_t)k * L + l] = 0;
        }
        if (out->all_dist) out->all_dist[k] = d;
        if (n_ok > 0) {
            /* strict <, first index wins: NaiveControlSampler.cpp:55 (double), RRT.cpp:466-467 (float) */
            int better;
            if (!in->dest) {
                better = best_k < 0;
            } else if (in->compare_f32) {
                better = (float)d < best_f;
            } else {
                better = d < best_d;
            }
            if (better) {
                best_k = k;
                best_d = d;
                best_f = (float)d;
                best_nok = n_ok;
                best_goal_step = goal_step;
                for (l = 0; l < L; ++l) {
                    if (out->best_states)
                        memcpy(out->best_states + (size_t)l * 3 * B, chain_states[l], sizeof(float) * 3 * (size_t)B);
                    if (out->best_controls)
                        for (i = 0; i < 5; ++i) out->best_controls[l * 5 + i] = chain_ctrl[l][i];
                    if (out->best_flags) out->best_flags[l] = chain_flags[l];
                }
            }
        }
    }
    if (out->best) {
        out->best->k = best_k < 0 ? -1 : best_k + in->k_offset;
        out->best->n_ok = best_nok;
        out->best->goal_step = best_goal_step;
        out->best->tree_index = -1;
        out->best->distance = best_k < 0 ? INFINITY : best_d;
    }
    return MPS_OK;
}

/* ------------------------------------------------------------------------- */
/* nearest neighbour: exact linear scan                                       */
/* ------------------------------------------------------------------------- */

int64_t orc_nn_linear(const mps_scene* sc, const float* states, const int32_t* slice_ids, int64_t n,
                      const float* query, const float* weights, uint32_t mask, int32_t slice_filter,
                      double* best_dist)
{
    int64_t best = -1, i;
    double bd = INFINITY;
    size_t stride = 3 * (size_t)sc->n_bodies;
    for (i = 0; i < n; ++i) {
        double d;
        if (slice_filter >= 0 && (!slice_ids || slice_ids[i] != slice_filter)) continue;
        d = orc_distance(sc, states + (size_t)i * stride, query, weights, mask);
        if (best < 0 || d < bd) { /* strict <: ties keep the lowest index */
            bd = d;
            best = i;
        }
    }
    if (best_dist) *best_dist = bd;
    return best;
}
