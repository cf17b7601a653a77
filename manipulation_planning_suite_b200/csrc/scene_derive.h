// scene_derive.h — mps_scene (ABI) -> DevScene (what the kernels read): derived constants and the pair table.
// Plain host code, shared by abi.cu and the test-side SIMT emulator build of the kernels (tests/emu).
#pragma once

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "mps_device.cuh"

static inline size_t mps_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static float bounding_radius(int shape, float hx, float hy)
{
    if (shape == MPS_SHAPE_DISC) return hx;
    return sqrtf(hx * hx + hy * hy);
}

// derived constants, same operation order as the physics spec
static void derive_scene(const mps_scene* sc, DevScene* d)
{
    memset(d, 0, sizeof(*d));
    const int B = sc->n_bodies, S = sc->n_statics;
    d->B = B;
    d->S = S;
    d->robot = sc->robot_id;
    d->vel_iters = sc->vel_iters;
    d->Mcap = sc->max_manifolds;
    d->strict = sc->strict_active_contacts;
    d->jam_steps = sc->jam_steps;
    d->slot_stride = (int)mps_align_up((size_t)(B + S), 4);
    d->dt = sc->dt;
    d->t_max = sc->t_max;
    d->slop = sc->slop;
    d->beta_dt = sc->baumgarte * (1.0f / sc->dt);
    d->max_bias = sc->max_bias_vel;
    d->pen_max = sc->pen_max;
    d->v_rest2 = sc->v_rest * sc->v_rest;
    d->w_rest = sc->w_rest;
    d->ref_tol = 0.1f * sc->slop;
    d->x_min = sc->x_min;
    d->x_max = sc->x_max;
    d->y_min = sc->y_min;
    d->y_max = sc->y_max;
    for (int i = 0; i < 3; ++i) d->accel[i] = sc->accel_limits[i];
    d->static_forbidden = sc->static_forbidden;
    d->pw = sc->position_weight;
    d->ow = sc->orientation_weight;
    for (int i = 0; i < B; ++i) {
        d->shape[i] = sc->shape[i];
        d->hx[i] = sc->hx[i];
        d->hy[i] = sc->hy[i];
        d->brad[i] = bounding_radius(sc->shape[i], sc->hx[i], sc->hy[i]);
        d->mass[i] = sc->mass[i];
        d->inertia[i] = sc->inertia[i];
        d->mu_c[i] = sc->mu_contact[i];
        d->pair_forbidden[i] = sc->pair_forbidden[i];
        if (i == sc->robot_id) {
            d->inv_m[i] = 0.0f;
            d->inv_i[i] = 0.0f;
            d->lin_cap[i] = 0.0f;
            d->ang_cap[i] = 0.0f;
        } else {
            float mg = sc->mass[i] * sc->gravity;
            float f = sc->mu_ground[i] * mg;
            d->inv_m[i] = 1.0f / sc->mass[i];
            d->inv_i[i] = 1.0f / sc->inertia[i];
            d->lin_cap[i] = sc->dt * f;
            d->ang_cap[i] = sc->dt * (f * sc->fr_radius[i]);
        }
    }
    for (int k = 0; k < S; ++k) {
        d->s_shape[k] = sc->s_shape[k];
        d->s_x[k] = sc->s_x[k];
        d->s_y[k] = sc->s_y[k];
        mps_sincos(sc->s_theta[k], &d->s_s[k], &d->s_c[k]);
        d->s_hx[k] = sc->s_hx[k];
        d->s_hy[k] = sc->s_hy[k];
        d->s_brad[k] = bounding_radius(sc->s_shape[k], sc->s_hx[k], sc->s_hy[k]);
        d->s_mu[k] = sc->s_mu_contact[k];
    }
    // The pair table is laid out in the solver's canonical order — colour r < B holds the body pairs with
    // (i + j) mod B == r, colour B + k the pairs (i, static k), ascending i inside a colour — so the manifold
    // list the detection compacts from it is already in solve order (rollout.cu: round scheduling).
    int P = 0;
    for (int r = 0; r < B; ++r)
        for (int i = 0; i < B; ++i) {
            int j = r - i;
            if (j < 0) j += B;
            if (j > i) d->pairs[P++] = (uint16_t)(i | (j << 8));
        }
    for (int k = 0; k < S; ++k)
        for (int i = 0; i < B; ++i) d->pairs[P++] = (uint16_t)(i | ((B + k) << 8));
    d->P = P;
    if (d->Mcap > P) d->Mcap = P;  // no more manifolds than pairs: same overflow behaviour, less shared memory
    const char* fg = getenv("MPS_FORCE_GENERAL_SWEEP");
    d->force_general = (fg && fg[0] == '1') ? 1 : 0;
}

