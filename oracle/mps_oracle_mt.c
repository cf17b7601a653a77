/*
 * mps_oracle_mt.c — pthread fan-out of the CPU oracle for the host-CPU baseline
 * ("one world per thread").  TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * The reference itself is single-threaded (world mutex held for a whole rollout,
 * SimEnvStatePropagator.cpp:57; global RNG, util/Random.cpp:9-16); the multi-thread
 * figure is reported next to the 1-thread one for transparency.
 */
#include "mps_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct ext_job {
    const mps_scene* sc;
    mps_extend_in in;
    mps_extend_out out;
    mps_extend_best best;
    float* best_states;
    float* best_controls;
    uint32_t* best_flags;
    int status;
} ext_job;

static void* ext_worker(void* p)
{
    ext_job* j = (ext_job*)p;
    j->status = orc_extend(j->sc, &j->in, &j->out);
    return 0;
}

/* K chains split into contiguous slices, one per thread; winners merged in slice order with the
 * same strict-< rule, so the result equals the sequential orc_extend. */
int orc_extend_mt(const mps_scene* sc, const mps_extend_in* in, const mps_extend_out* out, int n_threads)
{
    int B = sc->n_bodies, K = in->K, L = in->L, t, l, i;
    int nt = n_threads < 1 ? 1 : n_threads;
    ext_job* jobs;
    pthread_t* th;
    int best_t = -1;
    double best_d = DBL_MAX;
    float best_f = FLT_MAX;
    if (nt > K) nt = K > 0 ? K : 1;
    jobs = (ext_job*)calloc((size_t)nt, sizeof(ext_job));
    th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    for (t = 0; t < nt; ++t) {
        int k0 = (int)((long)K * t / nt), k1 = (int)((long)K * (t + 1) / nt);
        ext_job* j = &jobs[t];
        j->sc = sc;
        j->in = *in;
        j->in.K = k1 - k0;
        j->in.controls = in->controls + (size_t)k0 * L * 5;
        j->in.n_ctrl = in->n_ctrl ? in->n_ctrl + k0 : 0;
        j->in.starts = in->starts ? in->starts + (size_t)k0 * 3 * B : 0;
        j->in.k_offset = in->k_offset + k0;
        memset(&j->out, 0, sizeof(j->out));
        j->best_states = (float*)calloc((size_t)L * 3 * B, sizeof(float));
        j->best_controls = (float*)calloc((size_t)L * 5, sizeof(float));
        j->best_flags = (uint32_t*)calloc((size_t)L, sizeof(uint32_t));
        j->out.best = &j->best;
        j->out.best_states = j->best_states;
        j->out.best_controls = j->best_controls;
        j->out.best_flags = j->best_flags;
        if (out->all_states) j->out.all_states = out->all_states + (size_t)k0 * L * 3 * B;
        if (out->all_rest) j->out.all_rest = out->all_rest + (size_t)k0 * L;
        if (out->all_flags) j->out.all_flags = out->all_flags + (size_t)k0 * L;
        if (out->all_dist) j->out.all_dist = out->all_dist + k0;
        if (out->all_steps) j->out.all_steps = out->all_steps + (size_t)k0 * L;
        pthread_create(&th[t], 0, ext_worker, j);
    }
    for (t = 0; t < nt; ++t) pthread_join(th[t], 0);
    for (t = 0; t < nt; ++t) {
        ext_job* j = &jobs[t];
        if (j->best.k >= 0) {
            int better;
            if (!in->dest)
                better = best_t < 0;
            else if (in->compare_f32)
                better = (float)j->best.distance < best_f;
            else
                better = j->best.distance < best_d;
            if (better) {
                best_t = t;
                best_d = j->best.distance;
                best_f = (float)j->best.distance;
            }
        }
    }
    if (out->best) {
        if (best_t >= 0) {
            *out->best = jobs[best_t].best;
        } else {
            out->best->k = -1;
            out->best->n_ok = 0;
            out->best->goal_step = -1;
            out->best->tree_index = -1;
            out->best->distance = INFINITY;
        }
    }
    if (best_t >= 0) {
        ext_job* j = &jobs[best_t];
        for (l = 0; l < L; ++l) {
            if (out->best_states)
                memcpy(out->best_states + (size_t)l * 3 * B, j->best_states + (size_t)l * 3 * B,
                       sizeof(float) * 3 * (size_t)B);
            if (out->best_controls)
                for (i = 0; i < 5; ++i) out->best_controls[l * 5 + i] = j->best_controls[l * 5 + i];
            if (out->best_flags) out->best_flags[l] = j->best_flags[l];
        }
    }
    for (t = 0; t < nt; ++t) {
        free(jobs[t].best_states);
        free(jobs[t].best_controls);
        free(jobs[t].best_flags);
    }
    free(jobs);
    free(th);
    return MPS_OK;
}

typedef struct nn_job {
    const mps_scene* sc;
    const float* states;
    const int32_t* slice_ids;
    int64_t n;
    const float* queries;
    const float* weights;
    const uint32_t* masks;
    const int32_t* filters;
    int q0, q1;
    int64_t* indices;
    double* dists;
} nn_job;

static void* nn_worker(void* p)
{
    nn_job* j = (nn_job*)p;
    int q;
    size_t stride = 3 * (size_t)j->sc->n_bodies;
    for (q = j->q0; q < j->q1; ++q) {
        uint32_t mask = j->masks ? j->masks[q] : 0xffffffffu;
        int32_t filt = j->filters ? j->filters[q] : -1;
        double d;
        j->indices[q] = orc_nn_linear(j->sc, j->states, j->slice_ids, j->n, j->queries + (size_t)q * stride,
                                      j->weights, mask, filt, &d);
        if (j->dists) j->dists[q] = d;
    }
    return 0;
}

/* Q independent exact queries over the same node-major tree, split across threads */
int orc_nn_linear_batch_mt(const mps_scene* sc, const float* states, const int32_t* slice_ids, int64_t n,
                           const float* queries, int32_t Q, const float* weights, const uint32_t* masks,
                           const int32_t* filters, int64_t* indices, double* dists, int n_threads)
{
    int nt = n_threads < 1 ? 1 : n_threads, t;
    nn_job* jobs;
    pthread_t* th;
    if (nt > Q) nt = Q > 0 ? Q : 1;
    jobs = (nn_job*)calloc((size_t)nt, sizeof(nn_job));
    th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    for (t = 0; t < nt; ++t) {
        nn_job* j = &jobs[t];
        j->sc = sc;
        j->states = states;
        j->slice_ids = slice_ids;
        j->n = n;
        j->queries = queries;
        j->weights = weights;
        j->masks = masks;
        j->filters = filters;
        j->q0 = (int)((long)Q * t / nt);
        j->q1 = (int)((long)Q * (t + 1) / nt);
        j->indices = indices;
        j->dists = dists;
        pthread_create(&th[t], 0, nn_worker, j);
    }
    for (t = 0; t < nt; ++t) pthread_join(th[t], 0);
    free(jobs);
    free(th);
    return MPS_OK;
}
