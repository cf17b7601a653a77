// abi.cu — the extern "C" surface of libmps_b200.so (include/mps_b200.h): context, buffers,
// host<->device staging and kernel launches.  No arithmetic of the hot path lives here except the
// two host helpers mps_distance / mps_goal_distance (the reference's double expressions).
//
// No CPU fallback: every entry point that needs a device fails with MPS_ERR_NO_DEVICE / MPS_ERR_CUDA.
#include <cuda_runtime.h>

#include <cfloat>
#include <atomic>
#include <chrono>
#include <limits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mps_device.cuh"
#include "nn_scan.h"
#include "forest_step.h"
#include "peak.h"
#include "rollout.h"
#include "scene_derive.h"

namespace {
thread_local std::string g_last_error;
}

// words per chain written by the rollout kernel's cycle counters (8 in the -DMPS_PHASE_PROF profiling build)
#ifdef MPS_PHASE_PROF
#define MPS_CYCLE_WORDS 12
#else
#define MPS_CYCLE_WORDS 1
#endif

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct mps_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    mps_scene scene;
    DevScene hscene;
    DevScene* d_scene = nullptr;
    size_t scene_bytes = 0, warp_bytes = 0, warp_bytes_small = 0;
    int mcap_small = 0;
    RolloutTuning tuning = {};                 // launcher knobs, read from the environment at creation
    std::string err;
    unsigned long long launches = 0;

    // extend: device buffers
    DevBuf d_start, d_controls, d_nctrl, d_dest, d_weights, d_goals, d_ball;
    DevBuf d_inpack;  // all inputs of one mps_extend call, one host -> device copy
    DevBuf d_states, d_rest, d_flags, d_steps, d_dist, d_nok, d_goal_step;
    DevBuf d_record;  // packed winner record: best | states | controls | flags
    DevBuf d_work, d_counters, d_cycles, d_redo, d_order, d_gen_trunc;
    bool gen_ready = false;  // the controls of the uploaded extend were generated on the device
    // pinned staging
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    RolloutParams rp;
    SelectParams sp;
    bool extend_ready = false;
    int ext_K = 0, ext_L = 0;

    // tree
    TreeView tree;
    long long tree_n = 0;      // host mirror of the node count
    long long tree_upper = 0;  // upper bound including un-fetched device-side appends
    DevBuf d_tree_size;
    DevBuf d_tstage_states, d_tstage_parents, d_tstage_controls, d_tstage_slices;

    // nn
    DevBuf d_queries, d_masks, d_filters, d_nnw, d_nn_out, d_part_d, d_part_i, d_tickets;
    void* h_nn_out = nullptr;  // pinned: idx[Q] | dist[Q]
    unsigned long long* h_topk_rec = nullptr;  // pinned: the top-k scan's self-signed result records (nothing else writes here)
    size_t h_nn_out_bytes = 0;
    NNParams np;
    TopKParams topk;
    bool topk_ready = false;
    long long topk_fallbacks = 0;
    bool nn_ready = false;
    bool nn_direct = false;  // the pending query's result lands in h_nn_out without a copy
    unsigned int nn_seq = 0;  // sequence number of the last single query (completion flag in h_nn_out)
    char* h_valid = nullptr;                   // pinned: states in, results + signature out (mps_is_valid_batch, few states)
    unsigned long long* h_nn_parts = nullptr;  // pinned: per-CTA records of the host-folded single query
    int h_nn_parts_grid = 0;                   // grid of the last query that wrote them
    DevBuf d_topk_keys, d_topk_rej, d_topk_out;  // top-k: per-CTA lists, dropped minima, idx[32] | dist[32] | meta[2]
    int nn_Q = 0;

    // forest (multi-instance batches)
    TreeView forest;
    int f_ninst = 0;
    size_t f_seg = 0;
    bool f_reset = false;  // mps_forest_reset has run since the last reserve (every tree has its root)
    DevBuf d_fsizes, d_finst, d_fparents, d_fmasks, d_fbest, d_fbest_states, d_fbest_controls, d_fbest_flags, d_froots;

    // validity
    DevBuf d_vstates, d_valid, d_vflags;

    // the last CUDA call made for this context enqueued an NN scan (see MPS_CUDA_CHECK below and
    // mps_tree_nearest_launch): the next scan may overlap its tail
    bool nn_chain = false;

    // optional per-launch timing of the rollout kernel
    bool timing = false;
    std::vector<cudaEvent_t> ev0, ev1;
    int ev_count = 0;
};
#define MPS_EVENT_RING 512

// Every CUDA call made on behalf of a context goes through this macro; each one ends the chain of back-to-back NN
// scans (nn_chain), so a scan is only ever allowed to overlap the previous kernel of the stream when NOTHING else -
// no append, no upload, no allocation, no other launch - was issued for the context in between.
#undef MPS_CUDA_CHECK
#define MPS_CUDA_CHECK(expr, ctx_)                                                        \
    do {                                                                                  \
        mps_ctx* chk_ctx__ = (ctx_);                                                      \
        if (chk_ctx__) chk_ctx__->nn_chain = false;                                       \
        cudaError_t err__ = (expr);                                                       \
        if (err__ != cudaSuccess) return mps_fail_cuda(chk_ctx__, err__, #expr, __LINE__); \
    } while (0)

static int mps_fail(mps_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    g_last_error = msg;
    return code;
}

static int mps_fail_cuda(mps_ctx* ctx, cudaError_t e, const char* expr, int line)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at abi.cu:%d: %s", (int)e, cudaGetErrorString(e), line, expr);
    return mps_fail(ctx, e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MPS_ERR_NO_DEVICE : MPS_ERR_CUDA,
                    buf);
}

static int ensure(mps_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.bytes >= bytes) return MPS_OK;
    if (b.p) {
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaFree(b.p), ctx);
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 2;
    MPS_CUDA_CHECK(cudaMalloc(&b.p, want), ctx);
    b.bytes = want;
    return MPS_OK;
}

static int ensure_stage(mps_ctx* ctx, size_t bytes)
{
    if (ctx->h_stage_bytes >= bytes) return MPS_OK;
    if (ctx->h_stage) {
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaFreeHost(ctx->h_stage), ctx);
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
    }
    size_t want = bytes * 2 + 4096;
    MPS_CUDA_CHECK(cudaMallocHost(&ctx->h_stage, want), ctx);
    ctx->h_stage_bytes = want;
    return MPS_OK;
}

static void free_buf(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline size_t record_bytes(int B, int L)
{
    return align_up(sizeof(mps_extend_best) + sizeof(float) * (size_t)L * (3 * B + MPS_CONTROL_DIM) +
                        sizeof(uint32_t) * (size_t)L,
                    8);
}

// ---- scene ------------------------------------------------------------------

static int validate_scene(const mps_scene* sc, std::string& why)
{
    if (!sc) { why = "scene is NULL"; return 0; }
    if (sc->n_bodies < 2 || sc->n_bodies > MPS_MAX_BODIES) { why = "n_bodies must be in [2, 32]"; return 0; }
    if (sc->n_statics < 0 || sc->n_statics > MPS_MAX_STATICS) { why = "n_statics must be in [0, 16]"; return 0; }
    if (sc->robot_id < 0 || sc->robot_id >= sc->n_bodies) { why = "robot_id out of range"; return 0; }
    if (sc->vel_iters < 1 || sc->vel_iters > 64) { why = "vel_iters must be in [1, 64]"; return 0; }
    if (!(sc->dt > 0.0f)) { why = "dt must be > 0"; return 0; }
    if (sc->max_manifolds < 1 || sc->max_manifolds > MPS_MAX_MANIFOLDS) { why = "max_manifolds must be in [1, 64]"; return 0; }
    for (int i = 0; i < sc->n_bodies; ++i) {
        if (sc->shape[i] != MPS_SHAPE_BOX && sc->shape[i] != MPS_SHAPE_DISC) { why = "unknown body shape"; return 0; }
        if (!(sc->hx[i] > 0.0f)) { why = "body half extent must be > 0"; return 0; }
        if (sc->shape[i] == MPS_SHAPE_BOX && !(sc->hy[i] > 0.0f)) { why = "box hy must be > 0"; return 0; }
        if (i != sc->robot_id && (!(sc->mass[i] > 0.0f) || !(sc->inertia[i] > 0.0f))) { why = "mass/inertia must be > 0"; return 0; }
    }
    for (int i = 0; i < sc->n_statics; ++i) {
        if (sc->s_shape[i] != MPS_SHAPE_BOX && sc->s_shape[i] != MPS_SHAPE_DISC) { why = "unknown static shape"; return 0; }
        if (!(sc->s_hx[i] > 0.0f)) { why = "static half extent must be > 0"; return 0; }
    }
    for (int i = 0; i < 3; ++i)
        if (!(sc->accel_limits[i] > 0.0f)) { why = "accel_limits must be > 0"; return 0; }
    return 1;
}

extern "C" {

int mps_abi_version(void) { return MPS_ABI_VERSION; }

const char* mps_last_error(const mps_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int mps_device_count(int* count)
{
    if (!count) return mps_fail(nullptr, MPS_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return mps_fail_cuda(nullptr, e, "cudaGetDeviceCount", __LINE__);
    }
    *count = n;
    return MPS_OK;
}

void mps_scene_defaults(mps_scene* sc)
{
    if (!sc) return;
    memset(sc, 0, sizeof(*sc));
    sc->vel_iters = 8;
    for (int i = 0; i < MPS_MAX_BODIES; ++i) {
        sc->shape[i] = MPS_SHAPE_BOX;
        sc->hx[i] = 0.06f;
        sc->hy[i] = 0.06f;
        sc->mass[i] = 0.35f;     // magnitudes the reference mentions: LearnedOracle.cpp:314-318
        sc->inertia[i] = 0.002f;
        sc->mu_ground[i] = 0.19f;
        sc->fr_radius[i] = 0.0459f;
        sc->mu_contact[i] = 0.3f;
    }
    for (int i = 0; i < MPS_MAX_STATICS; ++i) {
        sc->s_shape[i] = MPS_SHAPE_BOX;
        sc->s_hx[i] = 0.05f;
        sc->s_hy[i] = 0.05f;
        sc->s_mu_contact[i] = 0.3f;
    }
    // workspace bounds default to +-FLT_MAX (OraclePushPlanner.cpp:61-66)
    sc->x_min = -FLT_MAX;
    sc->x_max = FLT_MAX;
    sc->y_min = -FLT_MAX;
    sc->y_max = FLT_MAX;
    sc->accel_limits[0] = 2.0f;
    sc->accel_limits[1] = 2.0f;
    sc->accel_limits[2] = 4.0f;
    sc->position_weight = 1.0;     // SimEnvState.h:224-231
    sc->orientation_weight = 0.08;
    sc->dt = 0.01f;
    sc->t_max = 8.0f;              // SimEnvStatePropagator.h:31, OraclePushPlanner.cpp:52
    sc->gravity = 9.81f;
    sc->slop = 0.002f;
    sc->baumgarte = 0.2f;
    sc->max_bias_vel = 0.5f;
    sc->pen_max = 0.01f;
    sc->v_rest = 0.001f;
    sc->w_rest = 0.01f;
    sc->max_manifolds = MPS_MAX_MANIFOLDS;
    sc->strict_active_contacts = 0;
    sc->jam_steps = 25;
}

int mps_ctx_create_on_stream(const mps_scene* scene, int device, void* cuda_stream, mps_ctx** out)
{
    if (!out) return mps_fail(nullptr, MPS_ERR_ARG, "out is NULL");
    *out = nullptr;
    std::string why;
    if (!validate_scene(scene, why)) return mps_fail(nullptr, MPS_ERR_ARG, "invalid scene: " + why);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return mps_fail(nullptr, MPS_ERR_NO_DEVICE,
                        "no CUDA device: the extend step has no CPU fallback (cudaGetDeviceCount: " +
                            std::string(cudaGetErrorString(e)) + ")");
    if (device < 0 || device >= n) return mps_fail(nullptr, MPS_ERR_ARG, "device index out of range");
    mps_ctx* ctx = new mps_ctx();
    ctx->device = device;
    ctx->scene = *scene;
    derive_scene(scene, &ctx->hscene);
    int rc = MPS_OK;
    do {
        if ((e = cudaSetDevice(device)) != cudaSuccess) { rc = mps_fail_cuda(nullptr, e, "cudaSetDevice", __LINE__); break; }
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { rc = mps_fail_cuda(nullptr, e, "cudaGetDeviceProperties", __LINE__); break; }
        if (prop.major < 10) {
            rc = mps_fail(nullptr, MPS_ERR_NO_DEVICE, "device is not sm_100: libmps_b200 ships sm_100a code only");
            break;
        }
        ctx->num_sms = prop.multiProcessorCount;
        if (cuda_stream) {
            ctx->stream = (cudaStream_t)cuda_stream;
            ctx->own_stream = false;
        } else {
            if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = mps_fail_cuda(nullptr, e, "cudaStreamCreate", __LINE__); break; }
            ctx->own_stream = true;
        }
        if ((e = cudaMalloc(&ctx->d_scene, sizeof(DevScene))) != cudaSuccess) { rc = mps_fail_cuda(nullptr, e, "cudaMalloc scene", __LINE__); break; }
        if ((e = cudaMemcpy(ctx->d_scene, &ctx->hscene, sizeof(DevScene), cudaMemcpyHostToDevice)) != cudaSuccess) { rc = mps_fail_cuda(nullptr, e, "cudaMemcpy scene", __LINE__); break; }
        ctx->scene_bytes = mps_rollout_scene_bytes();
        ctx->warp_bytes = mps_rollout_warp_bytes(ctx->hscene);
        // manifold slots per chain of a packed launch (two / four chains per warp): enough for every contact
        // pattern seen in practice; a chain that needs more is run again by the one-chain-per-warp kernel
        ctx->mcap_small = 2 * ctx->hscene.B > 8 ? 2 * ctx->hscene.B : 8;
        if (const char* e = getenv("MPS_ROLLOUT_SMALL_CAP")) ctx->mcap_small = atoi(e);  // test hook
        if (ctx->mcap_small >= ctx->hscene.Mcap) ctx->mcap_small = 0;
        ctx->warp_bytes_small = ctx->mcap_small ? mps_rollout_group_bytes(ctx->hscene, ctx->mcap_small) : ctx->warp_bytes;
        ctx->tuning = mps_rollout_tuning_from_env();
        memset(&ctx->tree, 0, sizeof(ctx->tree));
        ctx->tree.B = scene->n_bodies;
        memset(&ctx->forest, 0, sizeof(ctx->forest));
        ctx->forest.B = scene->n_bodies;
    } while (0);
    if (rc != MPS_OK) {
        delete ctx;
        return rc;
    }
    *out = ctx;
    return MPS_OK;
}

int mps_ctx_create(const mps_scene* scene, int device, mps_ctx** out)
{
    return mps_ctx_create_on_stream(scene, device, nullptr, out);
}

void mps_ctx_destroy(mps_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_start, &ctx->d_controls, &ctx->d_nctrl, &ctx->d_dest, &ctx->d_weights, &ctx->d_goals,
                      &ctx->d_ball, &ctx->d_inpack, &ctx->d_states, &ctx->d_rest, &ctx->d_flags, &ctx->d_steps, &ctx->d_dist,
                      &ctx->d_nok, &ctx->d_goal_step, &ctx->d_record, &ctx->d_work, &ctx->d_cycles, &ctx->d_redo, &ctx->d_order, &ctx->d_counters, &ctx->d_tree_size, &ctx->d_tstage_states,
                      &ctx->d_tstage_parents, &ctx->d_tstage_controls, &ctx->d_tstage_slices, &ctx->d_queries,
                      &ctx->d_masks, &ctx->d_filters, &ctx->d_nnw, &ctx->d_nn_out, &ctx->d_part_d,
                      &ctx->d_part_i, &ctx->d_tickets, &ctx->d_vstates, &ctx->d_valid, &ctx->d_vflags, &ctx->d_fsizes,
                      &ctx->d_finst, &ctx->d_fparents, &ctx->d_fmasks, &ctx->d_fbest, &ctx->d_fbest_states,
                      &ctx->d_fbest_controls, &ctx->d_fbest_flags, &ctx->d_froots};
    for (DevBuf* b : bufs) free_buf(*b);
    if (ctx->tree.states) cudaFree(ctx->tree.states);
    if (ctx->tree.controls) cudaFree(ctx->tree.controls);
    if (ctx->tree.parent) cudaFree(ctx->tree.parent);
    if (ctx->tree.slice) cudaFree(ctx->tree.slice);
    if (ctx->forest.states) cudaFree(ctx->forest.states);
    if (ctx->forest.controls) cudaFree(ctx->forest.controls);
    if (ctx->forest.parent) cudaFree(ctx->forest.parent);
    if (ctx->forest.slice) cudaFree(ctx->forest.slice);
    if (ctx->d_scene) cudaFree(ctx->d_scene);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_nn_out) cudaFreeHost(ctx->h_nn_out);
    if (ctx->h_topk_rec) cudaFreeHost(ctx->h_topk_rec);
    if (ctx->h_nn_parts) cudaFreeHost(ctx->h_nn_parts);
    if (ctx->h_valid) cudaFreeHost(ctx->h_valid);
    for (cudaEvent_t e : ctx->ev0) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev1) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mps_ctx_synchronize(mps_ctx* ctx)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}

int mps_ctx_scene(const mps_ctx* ctx, mps_scene* out)
{
    if (!ctx || !out) return mps_fail(nullptr, MPS_ERR_ARG, "NULL argument");
    *out = ctx->scene;
    return MPS_OK;
}

// ---- host helpers: the reference's double expressions --------------------------------------------

int mps_distance(const mps_scene* sc, const float* a, const float* b, const float* weights, uint32_t mask, double* out)
{
    if (!sc || !a || !b || !out) return mps_fail(nullptr, MPS_ERR_ARG, "NULL argument");
    double dist = 0.0;
    for (int i = 0; i < sc->n_bodies; ++i) {
        if ((mask >> i) & 1u) {
            double cfg = mps_body_distance(a[3 * i], a[3 * i + 1], a[3 * i + 2], b[3 * i], b[3 * i + 1], b[3 * i + 2],
                                           sc->position_weight, sc->orientation_weight);
            double w = weights ? (double)weights[i] : 1.0;
            dist += w * cfg;
        }
    }
    *out = dist;
    return MPS_OK;
}

int mps_goal_distance(const mps_scene* sc, const float* state, const mps_goal* goals, int32_t n_goals, double* out)
{
    if (!sc || !state || !out || (n_goals > 0 && !goals)) return mps_fail(nullptr, MPS_ERR_ARG, "NULL argument");
    double distance = 0.0;
    for (int g = 0; g < n_goals; ++g) {
        if (goals[g].body < 0 || goals[g].body >= sc->n_bodies) return mps_fail(nullptr, MPS_ERR_ARG, "goal body out of range");
        float fx = goals[g].x - state[3 * goals[g].body];
        float fy = goals[g].y - state[3 * goals[g].body + 1];
        double dc = sqrt((double)fx * (double)fx + (double)fy * (double)fy);
        double e = dc - (double)goals[g].tol;
        distance += e > 0.0 ? e : 0.0;
    }
    *out = distance;
    return MPS_OK;
}

// ---- tree ---------------------------------------------------------------------------------------

static int tree_alloc(mps_ctx* ctx, TreeView& t, size_t cap)
{
    const int B = ctx->scene.n_bodies;
    memset(&t, 0, sizeof(t));
    t.B = B;
    t.cap = cap;
    MPS_CUDA_CHECK(cudaMalloc(&t.states, sizeof(float) * 3 * B * cap), ctx);
    MPS_CUDA_CHECK(cudaMalloc(&t.controls, sizeof(float) * MPS_CONTROL_DIM * cap), ctx);
    MPS_CUDA_CHECK(cudaMalloc(&t.parent, sizeof(int32_t) * cap), ctx);
    MPS_CUDA_CHECK(cudaMalloc(&t.slice, sizeof(int32_t) * cap), ctx);
    return MPS_OK;
}

static int tree_ensure(mps_ctx* ctx, long long need)
{
    if (need < 0) return mps_fail(ctx, MPS_ERR_ARG, "negative tree capacity");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    int rc = ensure(ctx, ctx->d_tree_size, sizeof(long long));
    if (rc) return rc;
    if (!ctx->tree.states) MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_tree_size.p, 0, sizeof(long long), ctx->stream), ctx);
    if ((size_t)need <= ctx->tree.cap && ctx->tree.states) return MPS_OK;
    size_t cap = ctx->tree.cap ? ctx->tree.cap * 2 : 4096;
    if (cap < (size_t)need) cap = (size_t)need;
    cap = align_up(cap, 1024);
    TreeView nt;
    rc = tree_alloc(ctx, nt, cap);
    if (rc) return rc;
    if (ctx->tree.states) {
        MPS_CUDA_CHECK(mps_launch_tree_copy(ctx->tree, nt, ctx->tree_upper, ctx->stream), ctx);
        ctx->launches++;
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        cudaFree(ctx->tree.states);
        cudaFree(ctx->tree.controls);
        cudaFree(ctx->tree.parent);
        cudaFree(ctx->tree.slice);
    }
    ctx->tree = nt;  // launches read the tree pointers from ctx->tree each time
    return MPS_OK;
}

int mps_tree_reserve(mps_ctx* ctx, int64_t capacity)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    return tree_ensure(ctx, capacity);
}

int mps_tree_clear(mps_ctx* ctx)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    int rc = tree_ensure(ctx, 1);
    if (rc) return rc;
    ctx->tree_n = 0;
    ctx->tree_upper = 0;
    MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_tree_size.p, 0, sizeof(long long), ctx->stream), ctx);
    return MPS_OK;
}

static int tree_sync_size(mps_ctx* ctx)
{
    if (ctx->tree_upper == ctx->tree_n) return MPS_OK;
    long long n = 0;
    MPS_CUDA_CHECK(cudaMemcpyAsync(&n, ctx->d_tree_size.p, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    ctx->tree_n = n;
    ctx->tree_upper = n;
    return MPS_OK;
}

int mps_tree_size(mps_ctx* ctx, int64_t* n)
{
    if (!ctx || !n) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    if (!ctx->tree.states) {
        *n = 0;
        return MPS_OK;
    }
    int rc = tree_sync_size(ctx);
    if (rc) return rc;
    *n = ctx->tree_n;
    return MPS_OK;
}

int mps_tree_add_batch(mps_ctx* ctx, const float* states, const int32_t* parents, const float* controls,
                       const int32_t* slice_ids, int64_t n, int64_t* first_index)
{
    if (!ctx || !states || n < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_add_batch");
    const int B = ctx->scene.n_bodies;
    int rc = tree_ensure(ctx, 1);
    if (rc) return rc;
    if ((rc = tree_sync_size(ctx))) return rc;
    if ((rc = tree_ensure(ctx, ctx->tree_n + n))) return rc;
    if (n == 0) {
        if (first_index) *first_index = ctx->tree_n;
        return MPS_OK;
    }
    const size_t sb = sizeof(float) * 3 * B * (size_t)n, cb = sizeof(float) * MPS_CONTROL_DIM * (size_t)n,
                 ib = sizeof(int32_t) * (size_t)n;
    if ((rc = ensure(ctx, ctx->d_tstage_states, sb))) return rc;
    if ((rc = ensure(ctx, ctx->d_tstage_controls, cb))) return rc;
    if ((rc = ensure(ctx, ctx->d_tstage_parents, ib))) return rc;
    if ((rc = ensure(ctx, ctx->d_tstage_slices, ib))) return rc;
    if ((rc = ensure_stage(ctx, sb + cb + 2 * ib))) return rc;
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);  // staging buffer reuse
    char* h = (char*)ctx->h_stage;
    memcpy(h, states, sb);
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tstage_states.p, h, sb, cudaMemcpyHostToDevice, ctx->stream), ctx);
    h += sb;
    if (controls) {
        memcpy(h, controls, cb);
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tstage_controls.p, h, cb, cudaMemcpyHostToDevice, ctx->stream), ctx);
    }
    h += cb;
    if (parents) {
        memcpy(h, parents, ib);
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tstage_parents.p, h, ib, cudaMemcpyHostToDevice, ctx->stream), ctx);
    }
    h += ib;
    if (slice_ids) {
        memcpy(h, slice_ids, ib);
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tstage_slices.p, h, ib, cudaMemcpyHostToDevice, ctx->stream), ctx);
    }
    MPS_CUDA_CHECK(mps_launch_tree_append(ctx->tree, ctx->tree_n, n, (const float*)ctx->d_tstage_states.p,
                                          parents ? (const int32_t*)ctx->d_tstage_parents.p : nullptr,
                                          controls ? (const float*)ctx->d_tstage_controls.p : nullptr,
                                          slice_ids ? (const int32_t*)ctx->d_tstage_slices.p : nullptr,
                                          (long long*)ctx->d_tree_size.p, ctx->stream),
                   ctx);
    ctx->launches++;
    if (first_index) *first_index = ctx->tree_n;
    ctx->tree_n += n;
    ctx->tree_upper = ctx->tree_n;
    return MPS_OK;
}

int mps_tree_add(mps_ctx* ctx, const float* state, int32_t parent, const float* control, int32_t slice_id,
                 int64_t* index)
{
    return mps_tree_add_batch(ctx, state, &parent, control, &slice_id, 1, index);
}

int mps_tree_fill_random(mps_ctx* ctx, int64_t n, uint64_t seed, int32_t n_slices)
{
    if (!ctx || n < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_fill_random");
    int rc = tree_ensure(ctx, n > 0 ? n : 1);
    if (rc) return rc;
    float x0 = ctx->scene.x_min, x1 = ctx->scene.x_max, y0 = ctx->scene.y_min, y1 = ctx->scene.y_max;
    if (x0 < -1e6f) x0 = -1.0f;
    if (x1 > 1e6f) x1 = 1.0f;
    if (y0 < -1e6f) y0 = -1.0f;
    if (y1 > 1e6f) y1 = 1.0f;
    MPS_CUDA_CHECK(mps_launch_tree_fill(ctx->tree, n, seed, n_slices, x0, x1, y0, y1, (long long*)ctx->d_tree_size.p,
                                        ctx->stream),
                   ctx);
    ctx->launches++;
    ctx->tree_n = n;
    ctx->tree_upper = n;
    return MPS_OK;
}

int mps_tree_read_states(mps_ctx* ctx, int64_t first, int64_t n, float* states)
{
    if (!ctx || !states || first < 0 || n < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_read_states");
    int rc = tree_sync_size(ctx);
    if (rc) return rc;
    if (first + n > ctx->tree_n) return mps_fail(ctx, MPS_ERR_ARG, "range beyond the tree size");
    if (n == 0) return MPS_OK;
    const size_t sb = sizeof(float) * 3 * ctx->scene.n_bodies * (size_t)n;
    if ((rc = ensure(ctx, ctx->d_tstage_states, sb))) return rc;
    MPS_CUDA_CHECK(mps_launch_tree_gather(ctx->tree, first, n, (float*)ctx->d_tstage_states.p, ctx->stream), ctx);
    ctx->launches++;
    MPS_CUDA_CHECK(cudaMemcpyAsync(states, ctx->d_tstage_states.p, sb, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}

int mps_tree_get(mps_ctx* ctx, int64_t index, float* state, int32_t* parent, float* control, int32_t* slice_id)
{
    if (!ctx || index < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_get");
    int rc = tree_sync_size(ctx);
    if (rc) return rc;
    if (index >= ctx->tree_n) return mps_fail(ctx, MPS_ERR_ARG, "tree index out of range");
    const TreeView& t = ctx->tree;
    if (state) {
        MPS_CUDA_CHECK(cudaMemcpy2DAsync(state, sizeof(float), t.states + index, t.cap * sizeof(float), sizeof(float),
                                         3 * t.B, cudaMemcpyDeviceToHost, ctx->stream),
                       ctx);
    }
    if (control) {
        MPS_CUDA_CHECK(cudaMemcpy2DAsync(control, sizeof(float), t.controls + index, t.cap * sizeof(float),
                                         sizeof(float), MPS_CONTROL_DIM, cudaMemcpyDeviceToHost, ctx->stream),
                       ctx);
    }
    if (parent) MPS_CUDA_CHECK(cudaMemcpyAsync(parent, t.parent + index, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream), ctx);
    if (slice_id) MPS_CUDA_CHECK(cudaMemcpyAsync(slice_id, t.slice + index, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}

int mps_tree_backtrack(mps_ctx* ctx, int64_t index, int64_t* indices, int64_t max_len, int64_t* len)
{
    if (!ctx || !indices || !len || index < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_backtrack");
    int rc = tree_sync_size(ctx);
    if (rc) return rc;
    if (index >= ctx->tree_n) return mps_fail(ctx, MPS_ERR_ARG, "tree index out of range");
    std::vector<int32_t> parents((size_t)ctx->tree_n);
    MPS_CUDA_CHECK(cudaMemcpyAsync(parents.data(), ctx->tree.parent, sizeof(int32_t) * (size_t)ctx->tree_n,
                                   cudaMemcpyDeviceToHost, ctx->stream),
                   ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    std::vector<int64_t> rev;
    int64_t cur = index;
    while (cur >= 0) {
        rev.push_back(cur);
        if ((long long)rev.size() > ctx->tree_n) return mps_fail(ctx, MPS_ERR_STATE, "parent cycle in the tree");
        cur = parents[(size_t)cur];
    }
    if ((int64_t)rev.size() > max_len) {
        *len = (int64_t)rev.size();
        return mps_fail(ctx, MPS_ERR_CAPACITY, "path longer than max_len");
    }
    for (size_t i = 0; i < rev.size(); ++i) indices[i] = rev[rev.size() - 1 - i];
    *len = (int64_t)rev.size();
    return MPS_OK;
}

// ---- nearest neighbour ------------------------------------------------------------------------

int mps_tree_nearest_upload(mps_ctx* ctx, const float* queries, int32_t Q, const float* weights,
                            const uint32_t* masks, const int32_t* slice_filters)
{
    if (!ctx || !queries || Q < 1 || Q > 65535) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_nearest_upload");
    const int B = ctx->scene.n_bodies;
    int rc = tree_ensure(ctx, 1);
    if (rc) return rc;
    const size_t qb = sizeof(float) * 3 * B * (size_t)Q, mb = sizeof(uint32_t) * (size_t)Q, wb = sizeof(float) * B;
    if ((rc = ensure(ctx, ctx->d_queries, qb))) return rc;
    if ((rc = ensure(ctx, ctx->d_masks, mb))) return rc;
    if ((rc = ensure(ctx, ctx->d_filters, mb))) return rc;
    if ((rc = ensure(ctx, ctx->d_nnw, wb))) return rc;
    if ((rc = ensure(ctx, ctx->d_nn_out, 16 * (size_t)Q))) return rc;
    if (ctx->h_nn_out_bytes < 16 * (size_t)Q) {
        if (ctx->h_nn_out) {
            MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
            MPS_CUDA_CHECK(cudaFreeHost(ctx->h_nn_out), ctx);
            ctx->h_nn_out = nullptr;
        }
        ctx->h_nn_out_bytes = 16 * (size_t)(Q < 256 ? 256 : Q);
        MPS_CUDA_CHECK(cudaMallocHost(&ctx->h_nn_out, ctx->h_nn_out_bytes + 64), ctx);  // + the completion flag
        memset(ctx->h_nn_out, 0, ctx->h_nn_out_bytes + 64);
    }
    const int max_grid = ctx->num_sms * 16;
    if ((rc = ensure(ctx, ctx->d_part_d, sizeof(double) * (size_t)Q * max_grid))) return rc;
    if ((rc = ensure(ctx, ctx->d_part_i, sizeof(long long) * (size_t)Q * max_grid))) return rc;
    if (ctx->d_tickets.bytes < sizeof(unsigned) * (size_t)Q) {
        if ((rc = ensure(ctx, ctx->d_tickets, sizeof(unsigned) * (size_t)Q))) return rc;
        MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_tickets.p, 0, ctx->d_tickets.bytes, ctx->stream), ctx);
    }
    if (Q == 1) {
        // one query: it travels in the kernel parameters (filled below), nothing to copy
    } else if (qb + 2 * mb + wb <= 65536) {
        // small payloads: copy straight from the caller's memory (the driver stages pageable memory before
        // cudaMemcpyAsync returns), no staging buffer and no synchronise on the query path
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_queries.p, queries, qb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        if (masks) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_masks.p, masks, mb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        if (slice_filters) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_filters.p, slice_filters, mb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        if (weights) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_nnw.p, weights, wb, cudaMemcpyHostToDevice, ctx->stream), ctx);
    } else {
        if ((rc = ensure_stage(ctx, qb + 2 * mb + wb))) return rc;
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        char* h = (char*)ctx->h_stage;
        memcpy(h, queries, qb);
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_queries.p, h, qb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        h += qb;
        if (masks) {
            memcpy(h, masks, mb);
            MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_masks.p, h, mb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        }
        h += mb;
        if (slice_filters) {
            memcpy(h, slice_filters, mb);
            MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_filters.p, h, mb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        }
        h += mb;
        if (weights) {
            memcpy(h, weights, wb);
            MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_nnw.p, h, wb, cudaMemcpyHostToDevice, ctx->stream), ctx);
        }
    }
    NNParams& np = ctx->np;
    memset(&np, 0, sizeof(np));
    {
        // absolute slack of the fp32 candidate filter: the wrapped arc carries a few 1e-7 rad of absolute error
        // per body whatever the distance (nn_scan.cu); bounded with the largest weight and all B bodies
        float wmax = 1.0f;
        if (weights)
            for (int i = 0; i < B; ++i) wmax = fmaxf(wmax, fabsf(weights[i]));
        np.abs_eps = 1.0e-15f + 1.5e-6f * (float)ctx->scene.orientation_weight * wmax * (float)B;
    }
    np.queries = (const float*)ctx->d_queries.p;
    np.masks = masks ? (const uint32_t*)ctx->d_masks.p : nullptr;
    np.filters = slice_filters ? (const int32_t*)ctx->d_filters.p : nullptr;
    np.weights = weights ? (const float*)ctx->d_nnw.p : nullptr;
    np.pw = ctx->scene.position_weight;
    np.ow = ctx->scene.orientation_weight;
    np.Q = Q;
    np.out_idx = (long long*)ctx->d_nn_out.p;
    np.out_dist = (double*)((char*)ctx->d_nn_out.p + 8 * (size_t)Q);
    np.part_d = (double*)ctx->d_part_d.p;
    np.part_i = (long long*)ctx->d_part_i.p;
    np.tickets = (unsigned*)ctx->d_tickets.p;
    if (Q == 1) {
        np.inline_q = 1;
        np.inline_has_w = weights ? 1 : 0;
        np.inline_mask = masks ? masks[0] : 0xffffffffu;
        np.inline_filter = slice_filters ? slice_filters[0] : -1;
        memcpy(np.q_inline, queries, qb);
        if (weights) memcpy(np.w_inline, weights, wb);
    }
    ctx->nn_Q = Q;
    ctx->nn_ready = true;
    ctx->nn_direct = false;
    return MPS_OK;
}

int mps_tree_nearest_launch(mps_ctx* ctx)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    if (!ctx->nn_ready) return mps_fail(ctx, MPS_ERR_STATE, "mps_tree_nearest_upload must precede mps_tree_nearest_launch");
    int rc = tree_sync_size(ctx);
    if (rc) return rc;
    NNParams& np = ctx->np;
    np.tree = ctx->tree;
    np.n = ctx->tree_n;
    np.grid_x = mps_nn_grid_x(np.n, ctx->num_sms);
    // back-to-back scans of one context (nothing else issued for it since the last one): the new scan's streaming
    // pass may start while the previous scan's last CTAs finish (programmatic dependent launch, nn_scan.cu)
    np.overlap_prev = ctx->nn_chain ? 1 : 0;
    MPS_CUDA_CHECK(mps_launch_nn_scan(np, ctx->stream), ctx);
    ctx->nn_chain = true;
    ctx->launches++;
    return MPS_OK;
}

int mps_tree_nearest_fetch(mps_ctx* ctx, int64_t* indices, double* distances)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    if (!ctx->nn_ready) return mps_fail(ctx, MPS_ERR_STATE, "nothing to fetch");
    const int Q = ctx->nn_Q;
    // one D2H of idx[Q] | dist[Q] into pinned memory, one synchronise (the one-shot single query had the
    // kernel write the pinned block itself: synchronise only)
    if (!ctx->nn_direct)
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nn_out, ctx->d_nn_out.p, 16 * (size_t)Q, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    if (indices) memcpy(indices, ctx->h_nn_out, 8 * (size_t)Q);
    if (distances) memcpy(distances, (char*)ctx->h_nn_out + 8 * (size_t)Q, 8 * (size_t)Q);
    return MPS_OK;
}

int mps_tree_nearest_batch(mps_ctx* ctx, const float* queries, int32_t Q, const float* weights, const uint32_t* masks,
                           const int32_t* slice_filters, int64_t* indices, double* distances)
{
    int rc = mps_tree_nearest_upload(ctx, queries, Q, weights, masks, slice_filters);
    if (rc) return rc;
    if ((rc = mps_tree_nearest_launch(ctx))) return rc;
    return mps_tree_nearest_fetch(ctx, indices, distances);
}

// The host's half of the single query (nn_scan_kernel, host_parts): every CTA posts {distance bits, seq << 40 | index}
// as one 16-byte store into pinned memory; the host folds the records in CTA order as they arrive, with the kernel's
// own rule (smaller distance, then smaller index) - the same (double, index) minimum as the device-side hand-off,
// without its ticket, fences and last-CTA pass.  A 16-byte aligned store lies in one cache line, so a record whose
// second word carries this query's sequence number is complete.  Falls back to a stream synchronise if a record has
// not appeared after 20 ms (a failed launch never writes it).
static int nn_fold_parts(mps_ctx* ctx, int grid, unsigned int seq, int64_t* index, double* distance)
{
    const volatile unsigned long long* rec = ctx->h_nn_parts;
    const auto t0 = std::chrono::steady_clock::now();
    double bd = 0.0;
    long long bi = -1;
    bool synced = false;
    for (int c = 0; c < grid; ++c) {
        unsigned long long w1;
        unsigned spins = 0;
        while ((unsigned int)((w1 = rec[2 * c + 1]) >> 40) != seq) {
            if ((++spins & 0x3ff) == 0 && !synced &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.02) {
                MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
                synced = true;
                if ((unsigned int)(rec[2 * c + 1] >> 40) != seq)
                    return mps_fail(ctx, MPS_ERR_CUDA, "nearest-neighbour scan finished without a result");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        const unsigned long long w0 = rec[2 * c];
        const unsigned long long i40 = w1 & 0xffffffffffull;
        if (i40 == 0xffffffffffull) continue;  // this CTA saw no node (slice filter, ragged grid)
        double d;
        memcpy(&d, &w0, 8);
        const long long i = (long long)i40;
        if (bi < 0 || d < bd || (d == bd && i < bi)) {
            bd = d;
            bi = i;
        }
    }
    if (index) *index = bi;
    if (distance) *distance = bi < 0 ? std::numeric_limits<double>::infinity() : bd;
    return MPS_OK;
}

// waits for a sequence number a kernel writes into pinned host memory behind its results (a few hundred nanoseconds
// after the write, where a stream synchronise takes microseconds to return); falls back to the synchronise if the
// number has not appeared after 20 ms (a failed launch never writes it)
static int nn_wait_flag(mps_ctx* ctx, volatile unsigned int* flag, unsigned int seq)
{
    if (*flag != seq) {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while (*flag != seq) {
            if ((++spins & 0x3ff) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.02) {
                MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
                if (*flag != seq) return mps_fail(ctx, MPS_ERR_CUDA, "nearest-neighbour scan finished without a result");
                break;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return MPS_OK;
}

// waits for a self-signed 16-byte record {payload, seq << 40 | payload} a kernel stores into pinned host memory with ONE
// store (complete as soon as its second word carries the sequence number); same fallback as nn_wait_flag
static int nn_wait_record(mps_ctx* ctx, const volatile unsigned long long* rec, unsigned int seq)
{
    if ((unsigned int)(rec[1] >> 40) != seq) {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while ((unsigned int)(rec[1] >> 40) != seq) {
            if ((++spins & 0x3ff) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 0.02) {
                MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
                if ((unsigned int)(rec[1] >> 40) != seq)
                    return mps_fail(ctx, MPS_ERR_CUDA, "nearest-neighbour scan finished without a result");
                break;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return MPS_OK;
}

// pinned record block of the host-folded query and a fresh sequence number for a scan of `grid` CTAs
static int nn_parts_begin(mps_ctx* ctx, int grid)
{
    if (!ctx->h_nn_parts) {
        const size_t bytes = 16 * (size_t)ctx->num_sms * 16;  // the largest grid (mps_tree_nearest_upload)
        MPS_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_nn_parts, bytes), ctx);
        memset(ctx->h_nn_parts, 0, bytes);
    }
    if (grid != ctx->h_nn_parts_grid) {
        // no scan is in flight here (every host-folded query is waited for): records of another grid size must not be
        // mistaken for this one's when the sequence number comes round again
        memset(ctx->h_nn_parts, 0, 16 * (size_t)ctx->num_sms * 16);
        ctx->h_nn_parts_grid = grid;
    }
    do {
        ctx->nn_seq = (ctx->nn_seq + 1) & 0xffffffu;
    } while (ctx->nn_seq == 0);
    return MPS_OK;
}

int mps_tree_nearest(mps_ctx* ctx, const float* query, const float* weights, uint32_t mask, int32_t slice_filter,
                     int64_t* index, double* distance)
{
    // One query, one launch, no synchronise and no device-side hand-off: the query rides in the kernel parameters, every
    // CTA posts its result straight into pinned host memory and the host folds them (nn_fold_parts).
    int rc = mps_tree_nearest_upload(ctx, query, 1, weights, &mask, &slice_filter);
    if (rc) return rc;
    if ((rc = tree_sync_size(ctx))) return rc;
    const int grid = mps_nn_grid_x(ctx->tree_n, ctx->num_sms);
    if ((rc = nn_parts_begin(ctx, grid))) return rc;
    ctx->np.host_parts = ctx->h_nn_parts;
    ctx->np.host_seq = ctx->nn_seq;
    ctx->nn_direct = true;
    rc = mps_tree_nearest_launch(ctx);
    ctx->np.host_parts = nullptr;  // a re-launch of the cached query (mps_time_launches) takes the device-side hand-off
    ctx->np.out_idx = (long long*)ctx->h_nn_out;
    ctx->np.out_dist = (double*)((char*)ctx->h_nn_out + 8);
    if (rc) return rc;
    int64_t fi = -1;
    double fd = 0.0;
    if ((rc = nn_fold_parts(ctx, grid, ctx->nn_seq, &fi, &fd))) return rc;
    memcpy(ctx->h_nn_out, &fi, 8);  // where mps_tree_nearest_fetch looks
    memcpy((char*)ctx->h_nn_out + 8, &fd, 8);
    if (index) *index = fi;
    if (distance) *distance = fd;
    return MPS_OK;
}

// n single queries one after the other through mps_tree_nearest (host query in, (index, distance) out), timed by the
// host clock around the loop: the latency of the C entry point itself, without the caller's language binding
int mps_tree_nearest_timed(mps_ctx* ctx, const float* queries, int32_t n, const float* weights, uint32_t mask,
                           int32_t slice_filter, int64_t* indices, double* distances, double* seconds)
{
    if (!ctx || !queries || n < 1 || !seconds) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_tree_nearest_timed");
    const int B = ctx->scene.n_bodies;
    int64_t idx = -1;
    double d = 0.0;
    int rc = mps_tree_nearest(ctx, queries, weights, mask, slice_filter, &idx, &d);  // warm-up (buffers, first launch)
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) {
        rc = mps_tree_nearest(ctx, queries + (size_t)i * 3 * B, weights, mask, slice_filter, &idx, &d);
        if (rc) return rc;
        if (indices) indices[i] = idx;
        if (distances) distances[i] = d;
    }
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return MPS_OK;
}

// SliceBasedOracleRRT's two-level search without a host round trip between the levels (RRT.cpp:783-827 when the
// sample needs no oracle adjustment, i.e. the active object is the robot): scan 1 finds the nearest slice
// representative in `reps` (a second context whose node index is the slice id) under rep_mask, scan 2 - enqueued
// right behind it on this context's stream - finds the nearest node of the main tree under member_mask among the
// nodes of that slice, taking its slice filter from scan 1's result in device memory.  One wait for both.
int mps_tree_nearest_two_level(mps_ctx* ctx, mps_ctx* reps, const float* query, const float* weights, uint32_t rep_mask,
                               uint32_t member_mask, int64_t* rep_index, double* rep_distance, int64_t* index,
                               double* distance)
{
    if (!ctx || !reps || !query) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_tree_nearest_two_level");
    if (reps->device != ctx->device || reps->scene.n_bodies != ctx->scene.n_bodies)
        return mps_fail(ctx, MPS_ERR_ARG, "the two contexts must share the device and the scene");
    int rc;
    int32_t none = -1;
    // scan 1 on the representatives' tree, result in this ctx's device result block (slot 1)
    if ((rc = tree_sync_size(reps))) return rc;
    if (reps->stream != ctx->stream) MPS_CUDA_CHECK(cudaStreamSynchronize(reps->stream), ctx);  // its pending appends
    if ((rc = mps_tree_nearest_upload(ctx, query, 1, weights, &rep_mask, &none))) return rc;
    if ((rc = ensure(ctx, ctx->d_nn_out, 64))) return rc;
    NNParams p1 = ctx->np;
    p1.tree = reps->tree;
    p1.n = reps->tree_n;
    p1.grid_x = mps_nn_grid_x(p1.n, ctx->num_sms);
    p1.out_idx = (long long*)ctx->d_nn_out.p;
    p1.out_dist = (double*)((char*)ctx->d_nn_out.p + 8);
    // scan 2 on the main tree, filter = scan 1's index (device memory); its CTAs post their results into pinned memory
    // and the host folds them (nn_fold_parts) - no stream synchronise.  Scan 1 mirrors its own result into pinned
    // memory as well: scan 2 starts only after scan 1 has completed (it waits for it before reading the filter), so
    // by the time scan 2's records arrive the mirror is there.
    if ((rc = tree_sync_size(ctx))) return rc;
    NNParams p2 = ctx->np;
    p2.tree = ctx->tree;
    p2.n = ctx->tree_n;
    p2.grid_x = mps_nn_grid_x(p2.n, ctx->num_sms);
    if ((rc = nn_parts_begin(ctx, p2.grid_x))) return rc;
    p1.mirror_idx = (long long*)((char*)ctx->h_nn_out + 16);
    p1.mirror_dist = (double*)((char*)ctx->h_nn_out + 24);
    // ... and signs it: the sequence number follows the mirror behind a system-scope fence, the host checks it below
    volatile unsigned int* flag1 = (volatile unsigned int*)((char*)ctx->h_nn_out + ctx->h_nn_out_bytes);
    p1.done_flag = flag1;
    p1.done_seq = ctx->nn_seq;
    MPS_CUDA_CHECK(mps_launch_nn_scan(p1, ctx->stream), ctx);
    p2.inline_mask = member_mask;
    p2.filter_from = p1.out_idx;
    p2.host_parts = ctx->h_nn_parts;
    p2.host_seq = ctx->nn_seq;
    MPS_CUDA_CHECK(mps_launch_nn_scan(p2, ctx->stream), ctx);
    ctx->launches += 2;
    ctx->nn_ready = false;
    int64_t fi = -1;
    double fd = 0.0;
    if ((rc = nn_fold_parts(ctx, p2.grid_x, ctx->nn_seq, &fi, &fd))) return rc;
    if ((rc = nn_wait_flag(ctx, flag1, ctx->nn_seq))) return rc;  // scan 1 finished before scan 2 started: normally long there
    if (index) *index = fi;
    if (distance) *distance = fd;
    if (rep_index) memcpy(rep_index, (char*)ctx->h_nn_out + 16, 8);
    if (rep_distance) memcpy(rep_distance, (char*)ctx->h_nn_out + 24, 8);
    return MPS_OK;
}

// ompl::NearestNeighbors<MotionPtr>::nearestK for the device tree: the k <= 32 nearest nodes in ascending order of
// (double distance, node index).  One streaming pass (nn_topk_kernel); the rare launch that reports an ambiguous
// result (near-equal distances crowding one warp: duplicate nodes) is redone as k successor scans in double.
int mps_tree_nearest_k(mps_ctx* ctx, const float* query, const float* weights, uint32_t mask, int32_t slice_filter,
                       int32_t k, int64_t* indices, double* distances, int32_t* n_found)
{
    if (!ctx || !query || !indices) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_tree_nearest_k");
    if (k < 1 || k > NN_TOPK_MAX) return mps_fail(ctx, MPS_ERR_ARG, "k must be in [1, 32]");
    int rc = mps_tree_nearest_upload(ctx, query, 1, weights, &mask, &slice_filter);
    if (rc) return rc;
    if ((rc = tree_sync_size(ctx))) return rc;
    if (ctx->tree_n >= (1ll << 32)) return mps_fail(ctx, MPS_ERR_CAPACITY, "top-k keys carry 32-bit node indices");
    const int max_grid = ctx->num_sms * 16;
    if ((rc = ensure(ctx, ctx->d_topk_keys, sizeof(unsigned long long) * (size_t)max_grid * NN_TOPK_MAX))) return rc;
    if ((rc = ensure(ctx, ctx->d_topk_rej, sizeof(unsigned int) * (size_t)max_grid))) return rc;
    if ((rc = ensure(ctx, ctx->d_topk_out, 16 * NN_TOPK_MAX + 16))) return rc;
    TopKParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.nn = ctx->np;
    tp.nn.tree = ctx->tree;
    tp.nn.n = ctx->tree_n;
    tp.nn.grid_x = mps_nn_grid_x(tp.nn.n, ctx->num_sms);
    tp.k = k;
    tp.part_keys = (unsigned long long*)ctx->d_topk_keys.p;
    tp.part_rej = (unsigned int*)ctx->d_topk_rej.p;
    // the last CTA writes the results straight into pinned host memory as self-signed 16-byte records (one store each:
    // a record that carries this query's sequence number is complete): no copy, no fence, no synchronise
    if (!ctx->h_topk_rec) {
        MPS_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_topk_rec, 16 * (NN_TOPK_MAX + 1)), ctx);
        memset(ctx->h_topk_rec, 0, 16 * (NN_TOPK_MAX + 1));
    }
    tp.out_rec = ctx->h_topk_rec;
    const unsigned int seq_before = ctx->nn_seq;
    do {
        ctx->nn_seq = (ctx->nn_seq + 1) & 0xffffffu;
    } while (ctx->nn_seq == 0);
    // the 24-bit sequence number came round: a record slot that no query has rewritten since (a larger k long ago) must
    // not be taken for this query's (nothing is in flight here: every top-k query is waited for)
    if (ctx->nn_seq < seq_before) memset(ctx->h_topk_rec, 0, 16 * (NN_TOPK_MAX + 1));
    tp.done_seq = ctx->nn_seq;
    ctx->topk = tp;
    ctx->topk_ready = true;
    MPS_CUDA_CHECK(mps_launch_nn_topk(tp, ctx->stream), ctx);
    ctx->launches++;
    const volatile unsigned long long* rec = ctx->h_topk_rec;
    if ((rc = nn_wait_record(ctx, rec + 2 * NN_TOPK_MAX, tp.done_seq))) return rc;
    const unsigned long long meta = rec[2 * NN_TOPK_MAX];
    int found = (int)(unsigned int)(meta & 0xffffffffull);
    const int status = (int)(unsigned int)(meta >> 32);
    static const bool force_fallback = getenv("MPS_NN_TOPK_FALLBACK") && getenv("MPS_NN_TOPK_FALLBACK")[0] == '1';
    if (status == 0 && !force_fallback) {
        for (int r = 0; r < found; ++r) {
            if ((rc = nn_wait_record(ctx, rec + 2 * r, tp.done_seq))) return rc;
            const unsigned long long w0 = rec[2 * r], w1 = rec[2 * r + 1];
            indices[r] = (int64_t)(w1 & 0xffffffffffull);
            if (distances) memcpy(&distances[r], &w0, 8);
        }
    } else {
        // successor scans: (d, i) strictly above the previous result, k times
        NextParams nx;
        memset(&nx, 0, sizeof(nx));
        nx.nn = tp.nn;
        nx.nn.out_idx = (long long*)ctx->d_nn_out.p;
        nx.nn.out_dist = (double*)((char*)ctx->d_nn_out.p + 8);
        nx.after_d = -1.0;
        nx.after_i = -1;
        found = 0;
        for (int j = 0; j < k; ++j) {
            MPS_CUDA_CHECK(mps_launch_nn_next(nx, ctx->stream), ctx);
            ctx->launches++;
            MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->h_nn_out, ctx->d_nn_out.p, 16, cudaMemcpyDeviceToHost, ctx->stream), ctx);
            MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
            long long bi;
            double bd;
            memcpy(&bi, ctx->h_nn_out, 8);
            memcpy(&bd, (char*)ctx->h_nn_out + 8, 8);
            if (bi < 0) break;
            indices[found] = bi;
            if (distances) distances[found] = bd;
            ++found;
            nx.after_d = bd;
            nx.after_i = bi;
        }
        ctx->topk_fallbacks++;
    }
    if (n_found) *n_found = found;
    ctx->nn_ready = false;
    return MPS_OK;
}

int mps_tree_nearest_k_fallbacks(mps_ctx* ctx, int64_t* count)
{
    if (!ctx || !count) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    *count = ctx->topk_fallbacks;
    return MPS_OK;
}

// ---- extend -----------------------------------------------------------------------------------

static int extend_upload_impl(mps_ctx* ctx, const mps_extend_in* in, const mps_hybrid_gen* gen);

int mps_extend_upload(mps_ctx* ctx, const mps_extend_in* in) { return extend_upload_impl(ctx, in, nullptr); }

int mps_extend_upload_generated(mps_ctx* ctx, const mps_extend_in* in, const mps_hybrid_gen* gen)
{
    if (!gen) return mps_fail(ctx, MPS_ERR_ARG, "gen is NULL");
    if (!ctx || !in) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_extend_upload_generated");
    if (!in->start || in->starts || !in->dest) return mps_fail(ctx, MPS_ERR_ARG, "generated chains start from `start` towards `dest`");
    const int B = ctx->scene.n_bodies;
    if (gen->fixed_object >= B || gen->n_targets < 0 || gen->n_targets > MPS_MAX_GOALS)
        return mps_fail(ctx, MPS_ERR_ARG, "fixed_object / n_targets out of range");
    for (int i = 0; i < gen->n_targets; ++i)
        if (gen->targets[i] < 0 || gen->targets[i] >= B) return mps_fail(ctx, MPS_ERR_ARG, "target body out of range");
    for (int i = 0; i < 3; ++i)
        if (!(gen->velocity_limits[i] > 0.0f)) return mps_fail(ctx, MPS_ERR_ARG, "velocity limits must be positive");
    if (!(gen->duration_max > 0.0f)) return mps_fail(ctx, MPS_ERR_ARG, "duration_max must be positive");
    return extend_upload_impl(ctx, in, gen);
}

static int extend_upload_impl(mps_ctx* ctx, const mps_extend_in* in, const mps_hybrid_gen* gen)
{
    if (!ctx || !in) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_extend_upload");
    if ((!in->start && !in->starts) || (!in->controls && !gen)) return mps_fail(ctx, MPS_ERR_ARG, "start / controls are NULL");
    if (in->K < 1 || in->L < 1 || in->L > MPS_MAX_CHAIN) return mps_fail(ctx, MPS_ERR_ARG, "K must be >= 1 and L in [1, 64]");
    if (in->n_goals < 0 || in->n_goals > MPS_MAX_GOALS) return mps_fail(ctx, MPS_ERR_ARG, "n_goals must be in [0, 8]");
    if (in->n_goals > 0 && !in->goals) return mps_fail(ctx, MPS_ERR_ARG, "goals is NULL");
    const int B = ctx->scene.n_bodies, K = in->K, L = in->L;
    for (int g = 0; g < in->n_goals; ++g)
        if (in->goals[g].body < 0 || in->goals[g].body >= B) return mps_fail(ctx, MPS_ERR_ARG, "goal body out of range");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    const size_t KL = (size_t)K * L;
    const size_t sb = sizeof(float) * 3 * B, cb = sizeof(float) * MPS_CONTROL_DIM * KL, nb = sizeof(int32_t) * (size_t)K,
                 wb = sizeof(float) * B, gb = sizeof(mps_goal) * MPS_MAX_GOALS;
    int rc;
    if ((rc = ensure(ctx, ctx->d_start, in->starts ? sb * (size_t)K : sb))) return rc;
    if ((rc = ensure(ctx, ctx->d_controls, cb))) return rc;
    if ((rc = ensure(ctx, ctx->d_nctrl, nb))) return rc;
    if ((rc = ensure(ctx, ctx->d_dest, sb))) return rc;
    if ((rc = ensure(ctx, ctx->d_weights, wb))) return rc;
    if ((rc = ensure(ctx, ctx->d_goals, gb))) return rc;
    if ((rc = ensure(ctx, ctx->d_ball, sb))) return rc;
    if ((rc = ensure(ctx, ctx->d_states, sizeof(float) * 3 * B * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_rest, sizeof(float) * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_flags, sizeof(uint32_t) * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_steps, sizeof(int32_t) * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_dist, sizeof(double) * (size_t)K))) return rc;
    if ((rc = ensure(ctx, ctx->d_nok, sizeof(int32_t) * (size_t)K))) return rc;
    if ((rc = ensure(ctx, ctx->d_goal_step, sizeof(int32_t) * (size_t)K))) return rc;
    if ((rc = ensure(ctx, ctx->d_record, record_bytes(B, L)))) return rc;
    if ((rc = ensure(ctx, ctx->d_work, sizeof(unsigned)))) return rc;
    if ((rc = ensure(ctx, ctx->d_cycles, sizeof(long long) * (size_t)K * MPS_CYCLE_WORDS))) return rc;
    if ((rc = ensure(ctx, ctx->d_redo, (size_t)K))) return rc;
    if ((rc = ensure(ctx, ctx->d_order, sizeof(int32_t) * (size_t)K))) return rc;
    if (!ctx->d_counters.p) {
        if ((rc = ensure(ctx, ctx->d_counters, sizeof(unsigned long long) * 8))) return rc;
        MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long) * 8, ctx->stream), ctx);
    }
    if (in->append_to_tree) {
        if ((rc = tree_ensure(ctx, ctx->tree_upper + L))) return rc;
    }
    // host -> pinned staging -> device: every input of the call is packed into one block (16-byte aligned
    // pieces) and travels in ONE copy; the kernels read the pieces where they land in the device block
    const size_t pack_bytes = 3 * sb + cb + nb + wb + gb + (in->starts ? sb * (size_t)K + 64 : 0) + 256;
    // (the staging block also receives the winner record in mps_extend_fetch)
    if ((rc = ensure_stage(ctx, pack_bytes > record_bytes(B, L) ? pack_bytes : record_bytes(B, L)))) return rc;
    if ((rc = ensure(ctx, ctx->d_inpack, pack_bytes))) return rc;
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    char* const h0 = (char*)ctx->h_stage;
    char* h = h0;
    auto up = [&](const void* src, size_t bytes) -> const void* {
        memcpy(h, src, bytes);
        const void* d = (const char*)ctx->d_inpack.p + (h - h0);
        h += align_up(bytes, 16);
        return d;
    };
    const void* dp_start = in->starts ? up(in->starts, sb * (size_t)K) : up(in->start, sb);
    const void* dp_controls = gen ? ctx->d_controls.p : up(in->controls, cb);
    const void* dp_nctrl = gen ? ctx->d_nctrl.p : (in->n_ctrl ? up(in->n_ctrl, nb) : nullptr);
    const void* dp_dest = in->dest ? up(in->dest, sb) : nullptr;
    const void* dp_weights = in->weights ? up(in->weights, wb) : nullptr;
    const void* dp_goals = in->n_goals > 0 ? up(in->goals, sizeof(mps_goal) * in->n_goals) : ctx->d_goals.p;
    const void* dp_ball = in->ball_center ? up(in->ball_center, sb) : nullptr;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_inpack.p, h0, (size_t)(h - h0), cudaMemcpyHostToDevice, ctx->stream), ctx);
    ctx->gen_ready = false;
    if (gen) {
        // the chains' controls are produced where they are consumed: one small kernel instead of the K x L x 20 B upload
        HybridGenParams gp;
        memset(&gp, 0, sizeof(gp));
        gp.gen = *gen;
        gp.K = K;
        gp.L = L;
        gp.B = B;
        gp.robot = ctx->scene.robot_id;
        for (int i = 0; i < 3; ++i) gp.accel[i] = ctx->scene.accel_limits[i];
        gp.start = (const float*)dp_start;
        gp.dest = (const float*)dp_dest;
        gp.controls = (float*)ctx->d_controls.p;
        gp.n_ctrl = (int32_t*)ctx->d_nctrl.p;
        if ((rc = ensure(ctx, ctx->d_gen_trunc, sizeof(unsigned int)))) return rc;
        MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_gen_trunc.p, 0, sizeof(unsigned int), ctx->stream), ctx);
        gp.truncated = (unsigned int*)ctx->d_gen_trunc.p;
        MPS_CUDA_CHECK(mps_launch_hybrid_controls(gp, ctx->stream), ctx);
        ctx->launches++;
        ctx->gen_ready = true;
    }

    RolloutParams& rp = ctx->rp;
    memset(&rp, 0, sizeof(rp));
    rp.scene = ctx->d_scene;
    rp.scene_bytes = ctx->scene_bytes;
    rp.warp_bytes = ctx->warp_bytes;
    rp.warp_bytes_small = ctx->warp_bytes_small;
    rp.mcap_small = ctx->mcap_small;
    rp.redo_only = 0;
    rp.redo = (uint8_t*)ctx->d_redo.p;
    rp.start = (const float*)dp_start;
    rp.start_stride = in->starts ? (size_t)3 * B : 0;
    rp.stop_at_goal = in->stop_at_goal;
    rp.controls = (const float*)dp_controls;
    rp.n_ctrl = (const int32_t*)dp_nctrl;
    rp.K = K;
    rp.L = L;
    rp.dest = (const float*)dp_dest;
    rp.weights = (const float*)dp_weights;
    rp.mask = in->mask;
    rp.goals = (const mps_goal*)dp_goals;
    rp.n_goals = in->n_goals;
    rp.ball_center = (const float*)dp_ball;
    rp.ball_radius = in->ball_radius;
    rp.out_states = (float*)ctx->d_states.p;
    rp.out_rest = (float*)ctx->d_rest.p;
    rp.out_flags = (uint32_t*)ctx->d_flags.p;
    rp.out_steps = (int32_t*)ctx->d_steps.p;
    rp.out_dist = (double*)ctx->d_dist.p;
    rp.out_nok = (int32_t*)ctx->d_nok.p;
    rp.out_goal_step = (int32_t*)ctx->d_goal_step.p;
    rp.work_counter = (unsigned*)ctx->d_work.p;
    rp.counters = (unsigned long long*)ctx->d_counters.p;
    rp.out_cycles = (long long*)ctx->d_cycles.p;
    // longest chains first when the launch has more chains than one wave of warps can hide (tuning aid:
    // MPS_ROLLOUT_LPT=0 keeps the input order)
    static const bool lpt = !(getenv("MPS_ROLLOUT_LPT") && getenv("MPS_ROLLOUT_LPT")[0] == '0');
    rp.order = (lpt && K >= 64) ? (const int32_t*)ctx->d_order.p : nullptr;

    SelectParams& sp = ctx->sp;
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
    {
        char* rec = (char*)ctx->d_record.p;
        sp.best = (mps_extend_best*)rec;
        sp.best_states = (float*)(rec + sizeof(mps_extend_best));
        sp.best_controls = sp.best_states + (size_t)L * 3 * B;
        sp.best_flags = (uint32_t*)(sp.best_controls + (size_t)L * MPS_CONTROL_DIM);
    }
    sp.append = in->append_to_tree;
    sp.parent = in->parent;
    ctx->ext_K = K;
    ctx->ext_L = L;
    ctx->extend_ready = true;
    return MPS_OK;
}

int mps_extend_generated_controls(mps_ctx* ctx, float* controls, int32_t* n_ctrl, int32_t* n_truncated)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    if (!ctx->extend_ready || !ctx->gen_ready) return mps_fail(ctx, MPS_ERR_STATE, "no generated extend uploaded");
    const size_t KL = (size_t)ctx->ext_K * ctx->ext_L;
    unsigned int tr = 0;
    if (controls) MPS_CUDA_CHECK(cudaMemcpyAsync(controls, ctx->d_controls.p, sizeof(float) * MPS_CONTROL_DIM * KL, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    if (n_ctrl) MPS_CUDA_CHECK(cudaMemcpyAsync(n_ctrl, ctx->d_nctrl.p, sizeof(int32_t) * (size_t)ctx->ext_K, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaMemcpyAsync(&tr, ctx->d_gen_trunc.p, sizeof(tr), cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    if (n_truncated) *n_truncated = (int32_t)tr;
    return MPS_OK;
}

int mps_sample_controls(mps_ctx* ctx, uint64_t seed, int32_t iteration, int64_t gid, int32_t n, float velocity_limit,
                        float omega_limit, float duration_min, float duration_max, float* controls)
{
    if (!ctx || !controls || n < 1) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_sample_controls");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    int rc = ensure(ctx, ctx->d_vstates, sizeof(float) * MPS_CONTROL_DIM * (size_t)n);
    if (rc) return rc;
    MPS_CUDA_CHECK(mps_launch_sample_controls(seed, iteration, gid, n, velocity_limit, omega_limit, duration_min, duration_max,
                                              (float*)ctx->d_vstates.p, ctx->stream), ctx);
    ctx->launches++;
    MPS_CUDA_CHECK(cudaMemcpyAsync(controls, ctx->d_vstates.p, sizeof(float) * MPS_CONTROL_DIM * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}

int mps_extend_launch(mps_ctx* ctx)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "mps_extend_upload must precede mps_extend_launch");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    SelectParams& sp = ctx->sp;
    if (sp.append) {
        int rc = tree_ensure(ctx, ctx->tree_upper + ctx->ext_L);
        if (rc) return rc;
        sp.tree_size = (long long*)ctx->d_tree_size.p;
        sp.tree_cap = ctx->tree.cap;
        sp.tree_seg = ctx->tree.cap;
        sp.tree_states = ctx->tree.states;
        sp.tree_controls = ctx->tree.controls;
        sp.tree_parent = ctx->tree.parent;
        sp.tree_slice = ctx->tree.slice;
        ctx->tree_upper += ctx->ext_L;
    }
    if (ctx->rp.order) {
        MPS_CUDA_CHECK(mps_launch_chain_order(ctx->rp.controls, ctx->rp.n_ctrl, ctx->rp.K, ctx->rp.L, ctx->hscene.accel,
                                              ctx->hscene.dt, (int32_t*)ctx->d_order.p, ctx->stream), ctx);
        ctx->launches += 1;
    }
    const bool timed = ctx->timing && ctx->ev_count < MPS_EVENT_RING;
    if (timed) MPS_CUDA_CHECK(cudaEventRecord(ctx->ev0[ctx->ev_count], ctx->stream), ctx);
    MPS_CUDA_CHECK(mps_launch_rollout(ctx->rp, ctx->tuning, ctx->num_sms, ctx->scene.n_bodies, ctx->stream), ctx);
    if (timed) {
        MPS_CUDA_CHECK(cudaEventRecord(ctx->ev1[ctx->ev_count], ctx->stream), ctx);
        ctx->ev_count++;
    }
    MPS_CUDA_CHECK(mps_launch_select(sp, ctx->stream), ctx);
    ctx->launches += 2;
    return MPS_OK;
}

int mps_extend_fetch(mps_ctx* ctx, const mps_extend_out* out)
{
    if (!ctx || !out) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_extend_fetch");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "nothing to fetch");
    const int B = ctx->scene.n_bodies, K = ctx->ext_K, L = ctx->ext_L;
    const size_t KL = (size_t)K * L;
    cudaStream_t s = ctx->stream;
    // the winner record (best | states | controls | flags, contiguous on the device) comes back in one copy
    const bool want_rec = out->best || out->best_states || out->best_controls || out->best_flags;
    if (want_rec)
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->h_stage, ctx->d_record.p, record_bytes(B, L), cudaMemcpyDeviceToHost, s), ctx);
    if (out->all_states) MPS_CUDA_CHECK(cudaMemcpyAsync(out->all_states, ctx->d_states.p, sizeof(float) * 3 * B * KL, cudaMemcpyDeviceToHost, s), ctx);
    if (out->all_rest) MPS_CUDA_CHECK(cudaMemcpyAsync(out->all_rest, ctx->d_rest.p, sizeof(float) * KL, cudaMemcpyDeviceToHost, s), ctx);
    if (out->all_flags) MPS_CUDA_CHECK(cudaMemcpyAsync(out->all_flags, ctx->d_flags.p, sizeof(uint32_t) * KL, cudaMemcpyDeviceToHost, s), ctx);
    if (out->all_dist) MPS_CUDA_CHECK(cudaMemcpyAsync(out->all_dist, ctx->d_dist.p, sizeof(double) * (size_t)K, cudaMemcpyDeviceToHost, s), ctx);
    if (out->all_steps) MPS_CUDA_CHECK(cudaMemcpyAsync(out->all_steps, ctx->d_steps.p, sizeof(int32_t) * KL, cudaMemcpyDeviceToHost, s), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(s), ctx);
    if (want_rec) {
        const char* rec = (const char*)ctx->h_stage;
        const size_t o_states = sizeof(mps_extend_best), o_ctrl = o_states + sizeof(float) * 3 * B * L,
                     o_flags = o_ctrl + sizeof(float) * MPS_CONTROL_DIM * L;
        if (out->best) memcpy(out->best, rec, sizeof(mps_extend_best));
        if (out->best_states) memcpy(out->best_states, rec + o_states, sizeof(float) * 3 * B * L);
        if (out->best_controls) memcpy(out->best_controls, rec + o_ctrl, sizeof(float) * MPS_CONTROL_DIM * L);
        if (out->best_flags) memcpy(out->best_flags, rec + o_flags, sizeof(uint32_t) * L);
    }
    if (ctx->sp.append) return tree_sync_size(ctx);
    return MPS_OK;
}

int mps_extend(mps_ctx* ctx, const mps_extend_in* in, const mps_extend_out* out)
{
    int rc = mps_extend_upload(ctx, in);
    if (rc) return rc;
    if ((rc = mps_extend_launch(ctx))) return rc;
    return mps_extend_fetch(ctx, out);
}

// small pinned block shared by the synchronous low-latency calls (mps_is_valid_batch with few states, mps_propagate):
// 16 states of 32 bodies + results + a signature word
static int ensure_small_pinned(mps_ctx* ctx)
{
    if (!ctx->h_valid) MPS_CUDA_CHECK(cudaMallocHost((void**)&ctx->h_valid, 16 * 3 * MPS_MAX_BODIES * sizeof(float) + 256), ctx);
    return MPS_OK;
}

int mps_propagate(mps_ctx* ctx, const float* state_in, float* control, float* state_out, int32_t* ok, uint32_t* flags)
{
    // SimEnvStatePropagator::propagate as the reference calls it: one state, one control.  The rollout writes the
    // resulting state, rest time and flags straight into pinned host memory and the select kernel - the last work of
    // the extend - signs them; the host waits for the signature (three copies and a stream synchronise less: 151 ->
    // 124 us per call through ctypes, most of the rest being the rollout itself).
    if (!ctx || !state_in || !control || !state_out) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_propagate");
    mps_extend_in in;
    memset(&in, 0, sizeof(in));
    in.start = state_in;
    in.controls = control;
    in.K = 1;
    in.L = 1;
    int rc = mps_extend_upload(ctx, &in);
    if (rc) return rc;
    if ((rc = ensure_small_pinned(ctx))) return rc;
    const int B = ctx->scene.n_bodies;
    float* h_state = (float*)ctx->h_valid;
    float* h_rest = h_state + 3 * MPS_MAX_BODIES;
    uint32_t* h_flags = (uint32_t*)(h_rest + 1);
    volatile unsigned int* h_sig = (volatile unsigned int*)(h_rest + 2);
    ctx->rp.out_states = h_state;
    ctx->rp.out_rest = h_rest;
    ctx->rp.out_flags = h_flags;
    ctx->sp.states = h_state;
    ctx->sp.rest = h_rest;
    ctx->sp.flags = h_flags;
    do {
        ctx->nn_seq = (ctx->nn_seq + 1) & 0xffffffu;
    } while (ctx->nn_seq == 0);
    ctx->sp.done_flag = h_sig;
    ctx->sp.done_seq = ctx->nn_seq;
    *h_sig = 0u;
    rc = mps_extend_launch(ctx);
    ctx->extend_ready = false;  // the staged buffers do not hold this extend's outputs: nothing to fetch or relaunch
    if (rc) return rc;
    if ((rc = nn_wait_flag(ctx, h_sig, ctx->nn_seq))) return rc;
    memcpy(state_out, h_state, sizeof(float) * 3 * (size_t)B);
    control[4] = *h_rest;
    const uint32_t f = *h_flags;
    if (ok) *ok = (f & MPS_F_OK) ? 1 : 0;
    if (flags) *flags = f;
    return MPS_OK;
}

int mps_extend_chain_cycles(mps_ctx* ctx, int64_t* cycles)
{
    if (!ctx || !cycles) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "no extend uploaded");
    MPS_CUDA_CHECK(cudaMemcpyAsync(cycles, ctx->d_cycles.p, sizeof(long long) * (size_t)ctx->ext_K * MPS_CYCLE_WORDS, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}

int mps_extend_counters(mps_ctx* ctx, uint64_t counters[8])
{
    if (!ctx || !counters) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    memset(counters, 0, sizeof(uint64_t) * 8);
    if (ctx->d_counters.p) {
        MPS_CUDA_CHECK(cudaMemcpyAsync(counters, ctx->d_counters.p, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long) * 8, ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    }
    counters[4] = ctx->launches;
    return MPS_OK;
}

// ---- validity ---------------------------------------------------------------------------------

int mps_is_valid_batch(mps_ctx* ctx, const float* states, int32_t n, uint8_t* valid, uint32_t* flags)
{
    if (!ctx || !states || n < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_is_valid_batch");
    if (n == 0) return MPS_OK;
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    const size_t sb = sizeof(float) * 3 * ctx->scene.n_bodies * (size_t)n;
    int rc;
    if (n <= 16) {
        // A planner validates its samples one or a few at a time (RRT.cpp:171-190): one CTA reads the states from
        // pinned host memory, writes the results there and signs them; the host waits for the signature (43 -> 17 us
        // per call against copy in, launch, two copies out and a stream synchronise).
        const size_t vb = ((size_t)n + 15) & ~(size_t)15;
        // (its own pinned block: the staging buffer may still feed an asynchronous upload)
        if ((rc = ensure_small_pinned(ctx))) return rc;
        char* h = ctx->h_valid;
        memcpy(h, states, sb);
        ValidityParams vp;
        memset(&vp, 0, sizeof(vp));
        vp.scene = ctx->d_scene;
        vp.scene_bytes = ctx->scene_bytes;
        vp.warp_bytes = ctx->warp_bytes;
        vp.states = (const float*)h;
        vp.n = n;
        vp.valid = (uint8_t*)(h + sb);
        vp.flags = (uint32_t*)(h + sb + vb);
        vp.done_flag = (volatile unsigned int*)(h + sb + vb + sizeof(uint32_t) * (size_t)n);
        do {
            ctx->nn_seq = (ctx->nn_seq + 1) & 0xffffffu;
        } while (ctx->nn_seq == 0);
        vp.done_seq = ctx->nn_seq;
        *vp.done_flag = 0u;
        MPS_CUDA_CHECK(mps_launch_validity(vp, ctx->num_sms, ctx->stream), ctx);
        ctx->launches++;
        if ((rc = nn_wait_flag(ctx, vp.done_flag, vp.done_seq))) return rc;
        if (valid) memcpy(valid, h + sb, (size_t)n);
        if (flags) memcpy(flags, h + sb + vb, sizeof(uint32_t) * (size_t)n);
        return MPS_OK;
    }
    if ((rc = ensure(ctx, ctx->d_vstates, sb))) return rc;
    if ((rc = ensure(ctx, ctx->d_valid, (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->d_vflags, sizeof(uint32_t) * (size_t)n))) return rc;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_vstates.p, states, sb, cudaMemcpyHostToDevice, ctx->stream), ctx);
    ValidityParams vp;
    memset(&vp, 0, sizeof(vp));
    vp.scene = ctx->d_scene;
    vp.scene_bytes = ctx->scene_bytes;
    vp.warp_bytes = ctx->warp_bytes;
    vp.states = (const float*)ctx->d_vstates.p;
    vp.n = n;
    vp.valid = (uint8_t*)ctx->d_valid.p;
    vp.flags = (uint32_t*)ctx->d_vflags.p;
    MPS_CUDA_CHECK(mps_launch_validity(vp, ctx->num_sms, ctx->stream), ctx);
    ctx->launches++;
    if (valid) MPS_CUDA_CHECK(cudaMemcpyAsync(valid, ctx->d_valid.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    if (flags) MPS_CUDA_CHECK(cudaMemcpyAsync(flags, ctx->d_vflags.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}


// ---- forest: many independent planner instances batched per launch ----------------------------------

int mps_forest_reserve(mps_ctx* ctx, int32_t n_instances, int32_t cap_per_instance)
{
    if (!ctx || n_instances < 1 || cap_per_instance < 1) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_forest_reserve");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    if (ctx->forest.states) {
        cudaFree(ctx->forest.states);
        cudaFree(ctx->forest.controls);
        cudaFree(ctx->forest.parent);
        cudaFree(ctx->forest.slice);
        memset(&ctx->forest, 0, sizeof(ctx->forest));
    }
    const size_t seg = align_up((size_t)cap_per_instance, 32);
    int rc = tree_alloc(ctx, ctx->forest, seg * (size_t)n_instances);
    if (rc) return rc;
    ctx->f_ninst = n_instances;
    ctx->f_seg = seg;
    ctx->f_reset = false;
    if ((rc = ensure(ctx, ctx->d_fsizes, sizeof(long long) * (size_t)n_instances))) return rc;
    MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_fsizes.p, 0, sizeof(long long) * (size_t)n_instances, ctx->stream), ctx);
    return MPS_OK;
}

int mps_forest_reset(mps_ctx* ctx, const float* roots)
{
    if (!ctx || !roots) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_forest_reset");
    if (!ctx->forest.states) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reserve first");
    const int B = ctx->scene.n_bodies, n = ctx->f_ninst;
    const size_t sb = sizeof(float) * 3 * B * (size_t)n;
    int rc;
    if ((rc = ensure(ctx, ctx->d_froots, sb))) return rc;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_froots.p, roots, sb, cudaMemcpyHostToDevice, ctx->stream), ctx);
    MPS_CUDA_CHECK(mps_launch_forest_roots(ctx->forest, ctx->f_seg, n, (const float*)ctx->d_froots.p,
                                           (long long*)ctx->d_fsizes.p, ctx->stream),
                   ctx);
    ctx->launches++;
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    ctx->f_reset = true;
    return MPS_OK;
}

int mps_forest_sizes(mps_ctx* ctx, int64_t* sizes)
{
    if (!ctx || !sizes) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    if (!ctx->forest.states) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reserve first");
    MPS_CUDA_CHECK(cudaMemcpyAsync(sizes, ctx->d_fsizes.p, sizeof(long long) * (size_t)ctx->f_ninst, cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    return MPS_OK;
}

static int check_inst_ids(mps_ctx* ctx, const int32_t* ids, int M)
{
    for (int m = 0; m < M; ++m)
        if (ids[m] < 0 || ids[m] >= ctx->f_ninst) return mps_fail(ctx, MPS_ERR_ARG, "instance id out of range");
    return MPS_OK;
}

int mps_forest_nearest(mps_ctx* ctx, int32_t M, const int32_t* inst_ids, const float* queries, const float* weights,
                       const uint32_t* masks, int64_t* indices, double* distances)
{
    if (!ctx || M < 1 || M > 65535 || !inst_ids || !queries) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_forest_nearest");
    if (!ctx->forest.states) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reserve first");
    if (!ctx->f_reset) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reset first: the trees have no root yet");
    int rc = check_inst_ids(ctx, inst_ids, M);
    if (rc) return rc;
    // queries / masks / weights go through the plain query path's buffers
    if ((rc = mps_tree_nearest_upload(ctx, queries, M, weights, masks, nullptr))) return rc;
    if ((rc = ensure(ctx, ctx->d_finst, sizeof(int32_t) * (size_t)M))) return rc;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_finst.p, inst_ids, sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, ctx->stream), ctx);
    NNParams np = ctx->np;
    np.tree = ctx->forest;
    np.n = (long long)ctx->f_seg;  // grid sizing only; the kernel reads the real size of each segment
    np.inst_ids = (const int32_t*)ctx->d_finst.p;
    np.seg_sizes = (const long long*)ctx->d_fsizes.p;
    np.seg = ctx->f_seg;
    np.grid_x = mps_nn_grid_x((long long)ctx->f_seg, ctx->num_sms);
    int cap_grid = ctx->num_sms * 16 / (M < 1 ? 1 : M);
    if (cap_grid < 1) cap_grid = 1;
    if (np.grid_x > cap_grid) np.grid_x = cap_grid;  // many small trees: a few CTAs each
    MPS_CUDA_CHECK(mps_launch_nn_scan(np, ctx->stream), ctx);
    ctx->launches++;
    return mps_tree_nearest_fetch(ctx, indices, distances);
}

// device buffers of a forest extend of M instances x K chains x L controls
static int forest_ensure_buffers(mps_ctx* ctx, int M, int K, int L)
{
    const int B = ctx->scene.n_bodies;
    int rc;
    const size_t KT = (size_t)M * K, KL = KT * L;
    const size_t cb = sizeof(float) * MPS_CONTROL_DIM * KL, nb = sizeof(int32_t) * KT, db = sizeof(float) * 3 * B * (size_t)M,
                 mb = sizeof(int32_t) * (size_t)M, wb = sizeof(float) * B;
    if ((rc = ensure(ctx, ctx->d_controls, cb))) return rc;
    if ((rc = ensure(ctx, ctx->d_nctrl, nb))) return rc;
    if ((rc = ensure(ctx, ctx->d_dest, db))) return rc;
    if ((rc = ensure(ctx, ctx->d_weights, wb))) return rc;
    if ((rc = ensure(ctx, ctx->d_goals, sizeof(mps_goal) * MPS_MAX_GOALS))) return rc;
    if ((rc = ensure(ctx, ctx->d_finst, mb))) return rc;
    if ((rc = ensure(ctx, ctx->d_fparents, mb))) return rc;
    if ((rc = ensure(ctx, ctx->d_fmasks, mb))) return rc;
    if ((rc = ensure(ctx, ctx->d_states, sizeof(float) * 3 * B * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_rest, sizeof(float) * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_flags, sizeof(uint32_t) * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_steps, sizeof(int32_t) * KL))) return rc;
    if ((rc = ensure(ctx, ctx->d_dist, sizeof(double) * KT))) return rc;
    if ((rc = ensure(ctx, ctx->d_nok, sizeof(int32_t) * KT))) return rc;
    if ((rc = ensure(ctx, ctx->d_goal_step, sizeof(int32_t) * KT))) return rc;
    if ((rc = ensure(ctx, ctx->d_cycles, sizeof(long long) * KT * MPS_CYCLE_WORDS))) return rc;
    if ((rc = ensure(ctx, ctx->d_redo, (size_t)KT))) return rc;
    if ((rc = ensure(ctx, ctx->d_work, sizeof(unsigned)))) return rc;
    if (!ctx->d_counters.p) {
        if ((rc = ensure(ctx, ctx->d_counters, sizeof(unsigned long long) * 8))) return rc;
        MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned long long) * 8, ctx->stream), ctx);
    }
    if ((rc = ensure(ctx, ctx->d_fbest, sizeof(mps_extend_best) * (size_t)M))) return rc;
    if ((rc = ensure(ctx, ctx->d_fbest_states, sizeof(float) * 3 * B * L * (size_t)M))) return rc;
    if ((rc = ensure(ctx, ctx->d_fbest_controls, sizeof(float) * MPS_CONTROL_DIM * L * (size_t)M))) return rc;
    if ((rc = ensure(ctx, ctx->d_fbest_flags, sizeof(uint32_t) * L * (size_t)M))) return rc;
    return MPS_OK;
}

// rollout + select (+ append) of a forest extend whose inputs are already in the device buffers
static int forest_extend_launch(mps_ctx* ctx, int M, int K, int L, bool has_nctrl, bool has_dests, bool has_weights,
                                bool has_masks, int compare_f32, int n_goals, int append, mps_extend_best* best,
                                float* best_states, float* best_controls, uint32_t* best_flags)
{
    const int B = ctx->scene.n_bodies;
    const size_t KT = (size_t)M * K;
    cudaStream_t st = ctx->stream;
    RolloutParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.scene = ctx->d_scene;
    rp.scene_bytes = ctx->scene_bytes;
    rp.warp_bytes = ctx->warp_bytes;
    rp.warp_bytes_small = ctx->warp_bytes_small;
    rp.mcap_small = ctx->mcap_small;
    rp.redo_only = 0;
    rp.redo = (uint8_t*)ctx->d_redo.p;
    rp.controls = (const float*)ctx->d_controls.p;
    rp.n_ctrl = has_nctrl ? (const int32_t*)ctx->d_nctrl.p : nullptr;
    rp.K = (int)KT;
    rp.L = L;
    rp.dest = has_dests ? (const float*)ctx->d_dest.p : nullptr;
    rp.weights = has_weights ? (const float*)ctx->d_weights.p : nullptr;
    rp.mask = B < 32 ? (1u << B) - 1u : 0xffffffffu;
    rp.goals = (const mps_goal*)ctx->d_goals.p;
    rp.n_goals = n_goals;
    rp.out_states = (float*)ctx->d_states.p;
    rp.out_rest = (float*)ctx->d_rest.p;
    rp.out_flags = (uint32_t*)ctx->d_flags.p;
    rp.out_steps = (int32_t*)ctx->d_steps.p;
    rp.out_dist = (double*)ctx->d_dist.p;
    rp.out_nok = (int32_t*)ctx->d_nok.p;
    rp.out_goal_step = (int32_t*)ctx->d_goal_step.p;
    rp.work_counter = (unsigned*)ctx->d_work.p;
    rp.counters = (unsigned long long*)ctx->d_counters.p;
    rp.out_cycles = (long long*)ctx->d_cycles.p;
    rp.K_per = K;
    rp.inst_ids = (const int32_t*)ctx->d_finst.p;
    rp.parents = (const int32_t*)ctx->d_fparents.p;
    rp.f_states = ctx->forest.states;
    rp.f_cols = ctx->forest.cap;
    rp.f_seg = ctx->f_seg;
    rp.masks = has_masks ? (const uint32_t*)ctx->d_fmasks.p : nullptr;

    SelectParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.K = K;
    sp.L = L;
    sp.B = B;
    sp.M = M;
    sp.has_dest = has_dests ? 1 : 0;
    sp.compare_f32 = compare_f32;
    sp.dist = rp.out_dist;
    sp.nok = rp.out_nok;
    sp.goal_step = rp.out_goal_step;
    sp.states = rp.out_states;
    sp.rest = rp.out_rest;
    sp.flags = rp.out_flags;
    sp.controls = rp.controls;
    sp.best = (mps_extend_best*)ctx->d_fbest.p;
    sp.best_states = (float*)ctx->d_fbest_states.p;
    sp.best_controls = (float*)ctx->d_fbest_controls.p;
    sp.best_flags = (uint32_t*)ctx->d_fbest_flags.p;
    sp.append = append;
    sp.tree_size = (long long*)ctx->d_fsizes.p;
    sp.tree_cap = ctx->forest.cap;
    sp.tree_seg = ctx->f_seg;
    sp.tree_states = ctx->forest.states;
    sp.tree_controls = ctx->forest.controls;
    sp.tree_parent = ctx->forest.parent;
    sp.tree_slice = ctx->forest.slice;
    sp.inst_ids = rp.inst_ids;
    sp.parents = rp.parents;
    // longest chains first, as in the plain extend (the launch lasts as long as its last chains)
    static const bool lpt = !(getenv("MPS_ROLLOUT_LPT") && getenv("MPS_ROLLOUT_LPT")[0] == '0');
    if (lpt && KT >= 64) {
        int rc2 = ensure(ctx, ctx->d_order, sizeof(int32_t) * KT);
        if (rc2) return rc2;
        rp.order = (const int32_t*)ctx->d_order.p;
        MPS_CUDA_CHECK(mps_launch_chain_order(rp.controls, rp.n_ctrl, rp.K, rp.L, ctx->hscene.accel, ctx->hscene.dt,
                                              (int32_t*)ctx->d_order.p, st), ctx);
        ctx->launches += 1;
    }
    MPS_CUDA_CHECK(mps_launch_rollout(rp, ctx->tuning, ctx->num_sms, ctx->scene.n_bodies, st), ctx);
    MPS_CUDA_CHECK(mps_launch_select(sp, st), ctx);
    ctx->launches += 2;
    ctx->extend_ready = false;  // the plain extend's cached parameters no longer describe the buffers
    if (best) MPS_CUDA_CHECK(cudaMemcpyAsync(best, sp.best, sizeof(mps_extend_best) * (size_t)M, cudaMemcpyDeviceToHost, st), ctx);
    if (best_states) MPS_CUDA_CHECK(cudaMemcpyAsync(best_states, sp.best_states, sizeof(float) * 3 * B * L * (size_t)M, cudaMemcpyDeviceToHost, st), ctx);
    if (best_controls) MPS_CUDA_CHECK(cudaMemcpyAsync(best_controls, sp.best_controls, sizeof(float) * MPS_CONTROL_DIM * L * (size_t)M, cudaMemcpyDeviceToHost, st), ctx);
    if (best_flags) MPS_CUDA_CHECK(cudaMemcpyAsync(best_flags, sp.best_flags, sizeof(uint32_t) * L * (size_t)M, cudaMemcpyDeviceToHost, st), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(st), ctx);
    return MPS_OK;
}


int mps_forest_extend(mps_ctx* ctx, const mps_forest_extend_in* in, mps_extend_best* best, float* best_states,
                      float* best_controls, uint32_t* best_flags)
{
    if (!ctx || !in || !in->inst_ids || !in->parents || !in->controls) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_forest_extend");
    if (!ctx->forest.states) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reserve first");
    const int B = ctx->scene.n_bodies, M = in->M, K = in->K, L = in->L;
    if (M < 1 || K < 1 || L < 1 || L > MPS_MAX_CHAIN) return mps_fail(ctx, MPS_ERR_ARG, "M, K >= 1 and L in [1, 64]");
    if ((long long)M * K > (1ll << 24)) return mps_fail(ctx, MPS_ERR_ARG, "M * K too large");
    if (in->n_goals < 0 || in->n_goals > MPS_MAX_GOALS || (in->n_goals > 0 && !in->goals)) return mps_fail(ctx, MPS_ERR_ARG, "bad goals");
    if (!ctx->f_reset) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reset first: the trees have no root yet");
    for (int g = 0; g < in->n_goals; ++g)
        if (in->goals[g].body < 0 || in->goals[g].body >= B) return mps_fail(ctx, MPS_ERR_ARG, "goal body out of range");
    // the device reads the start state of instance m at column inst * seg + parents[m]: a parent outside its
    // segment would be an out-of-bounds read (nodes beyond the tree's size but inside the segment are the
    // caller's business, like an unknown node id in the reference)
    for (int m = 0; m < M; ++m)
        if (in->parents[m] < 0 || (size_t)in->parents[m] >= ctx->f_seg) return mps_fail(ctx, MPS_ERR_ARG, "parent node out of range");
    int rc = check_inst_ids(ctx, in->inst_ids, M);
    if (rc) return rc;
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    if ((rc = forest_ensure_buffers(ctx, M, K, L))) return rc;
    const size_t KT = (size_t)M * K, KL = KT * L;
    const size_t cb = sizeof(float) * MPS_CONTROL_DIM * KL, nb = sizeof(int32_t) * KT, db = sizeof(float) * 3 * B * (size_t)M,
                 mb = sizeof(int32_t) * (size_t)M, wb = sizeof(float) * B;
    cudaStream_t st = ctx->stream;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_controls.p, in->controls, cb, cudaMemcpyHostToDevice, st), ctx);
    if (in->n_ctrl) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_nctrl.p, in->n_ctrl, nb, cudaMemcpyHostToDevice, st), ctx);
    if (in->dests) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_dest.p, in->dests, db, cudaMemcpyHostToDevice, st), ctx);
    if (in->weights) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_weights.p, in->weights, wb, cudaMemcpyHostToDevice, st), ctx);
    if (in->n_goals > 0) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_goals.p, in->goals, sizeof(mps_goal) * in->n_goals, cudaMemcpyHostToDevice, st), ctx);
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_finst.p, in->inst_ids, mb, cudaMemcpyHostToDevice, st), ctx);
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_fparents.p, in->parents, mb, cudaMemcpyHostToDevice, st), ctx);
    if (in->masks) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_fmasks.p, in->masks, mb, cudaMemcpyHostToDevice, st), ctx);

    return forest_extend_launch(ctx, M, K, L, in->n_ctrl != nullptr, in->dests != nullptr, in->weights != nullptr,
                                in->masks != nullptr, in->compare_f32, in->n_goals, in->append_to_tree, best, best_states,
                                best_controls, best_flags);
}

int mps_forest_naive_step(mps_ctx* ctx, const mps_forest_naive_in* in, mps_extend_best* best, uint32_t* flags)
{
    if (!ctx || !in || !in->inst_ids) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_forest_naive_step");
    if (!ctx->forest.states) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reserve first");
    const int B = ctx->scene.n_bodies, M = in->M, K = in->K, C = in->C, L = 1;
    if (M < 1 || M > 65535 || K < 1 || C < 1) return mps_fail(ctx, MPS_ERR_ARG, "M in [1, 65535], K, C >= 1");
    if ((long long)M * K > (1ll << 24)) return mps_fail(ctx, MPS_ERR_ARG, "M * K too large");
    if (in->n_goals < 1 || in->n_goals > MPS_MAX_GOALS || !in->goals) return mps_fail(ctx, MPS_ERR_ARG, "bad goals");
    for (int g = 0; g < in->n_goals; ++g)
        if (in->goals[g].body < 0 || in->goals[g].body >= B) return mps_fail(ctx, MPS_ERR_ARG, "goal body out of range");
    if (!ctx->f_reset) return mps_fail(ctx, MPS_ERR_STATE, "mps_forest_reset first: the trees have no root yet");
    {
        // samples are drawn as min + (max - min) * u: the span must be finite (the scene default is +-FLT_MAX)
        const float sx = ctx->scene.x_max - ctx->scene.x_min, sy = ctx->scene.y_max - ctx->scene.y_min;
        if (!std::isfinite(sx) || !std::isfinite(sy) || sx <= 0.0f || sy <= 0.0f)
            return mps_fail(ctx, MPS_ERR_ARG, "the scene's workspace bounds must span a finite, positive range to sample states");
    }
    int rc = check_inst_ids(ctx, in->inst_ids, M);
    if (rc) return rc;
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    if ((rc = forest_ensure_buffers(ctx, M, K, L))) return rc;
    const size_t MC = (size_t)M * C;
    if ((rc = ensure(ctx, ctx->d_vstates, sizeof(float) * 3 * B * MC))) return rc;
    if ((rc = ensure(ctx, ctx->d_valid, MC))) return rc;
    if ((rc = ensure(ctx, ctx->d_vflags, sizeof(uint32_t) * MC))) return rc;
    // buffers of the nearest-neighbour query (as mps_tree_nearest_upload sets them up)
    if ((rc = ensure(ctx, ctx->d_nn_out, 16 * (size_t)M))) return rc;
    const int max_grid = ctx->num_sms * 16;
    if ((rc = ensure(ctx, ctx->d_part_d, sizeof(double) * (size_t)M * max_grid))) return rc;
    if ((rc = ensure(ctx, ctx->d_part_i, sizeof(long long) * (size_t)M * max_grid))) return rc;
    if (ctx->d_tickets.bytes < sizeof(unsigned) * (size_t)M) {
        if ((rc = ensure(ctx, ctx->d_tickets, sizeof(unsigned) * (size_t)M))) return rc;
        MPS_CUDA_CHECK(cudaMemsetAsync(ctx->d_tickets.p, 0, ctx->d_tickets.bytes, ctx->stream), ctx);
    }
    cudaStream_t st = ctx->stream;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_finst.p, in->inst_ids, sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, st), ctx);
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_goals.p, in->goals, sizeof(mps_goal) * in->n_goals, cudaMemcpyHostToDevice, st), ctx);
    if (in->weights) MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->d_weights.p, in->weights, sizeof(float) * B, cudaMemcpyHostToDevice, st), ctx);

    // 1. candidate states, active objects and controls
    NaiveSampleParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.M = M;
    sp.C = C;
    sp.K = K;
    sp.B = B;
    sp.inst_ids = (const int32_t*)ctx->d_finst.p;
    sp.instance_offset = in->instance_offset;
    sp.seed = in->seed;
    sp.iteration = in->iteration;
    sp.x_min = ctx->scene.x_min;
    sp.x_max = ctx->scene.x_max;
    sp.y_min = ctx->scene.y_min;
    sp.y_max = ctx->scene.y_max;
    sp.goal_bias = in->goal_bias;
    sp.target_bias = in->target_bias;
    sp.robot_bias = in->robot_bias;
    sp.robot = ctx->scene.robot_id;
    sp.goals = (const mps_goal*)ctx->d_goals.p;
    sp.n_goals = in->n_goals;
    sp.v_limit = in->velocity_limit;
    sp.w_limit = in->omega_limit;
    sp.dur_min = in->duration_min;
    sp.dur_max = in->duration_max;
    sp.cand = (float*)ctx->d_vstates.p;
    sp.masks = (uint32_t*)ctx->d_fmasks.p;
    sp.controls = (float*)ctx->d_controls.p;
    MPS_CUDA_CHECK(mps_launch_naive_sample(sp, st), ctx);
    // 2. validity of the candidates, first valid one becomes the sample (= dest of the extend, query of the NN)
    ValidityParams vp;
    memset(&vp, 0, sizeof(vp));
    vp.scene = ctx->d_scene;
    vp.scene_bytes = ctx->scene_bytes;
    vp.warp_bytes = ctx->warp_bytes;
    vp.states = (const float*)ctx->d_vstates.p;
    vp.n = (int)MC;
    vp.valid = (uint8_t*)ctx->d_valid.p;
    vp.flags = (uint32_t*)ctx->d_vflags.p;
    MPS_CUDA_CHECK(mps_launch_validity(vp, ctx->num_sms, st), ctx);
    MPS_CUDA_CHECK(mps_launch_naive_pick(M, C, B, (const float*)ctx->d_vstates.p, (const uint8_t*)ctx->d_valid.p,
                                         (float*)ctx->d_dest.p, st),
                   ctx);
    // 3. selectTreeNode: nearest node of the instance's own tree under the full metric
    NNParams np;
    memset(&np, 0, sizeof(np));
    np.tree = ctx->forest;
    np.n = (long long)ctx->f_seg;
    np.queries = (const float*)ctx->d_dest.p;
    np.weights = in->weights ? (const float*)ctx->d_weights.p : nullptr;
    np.pw = ctx->scene.position_weight;
    np.ow = ctx->scene.orientation_weight;
    np.Q = M;
    np.out_idx = (long long*)ctx->d_nn_out.p;
    np.out_dist = (double*)((char*)ctx->d_nn_out.p + 8 * (size_t)M);
    np.part_d = (double*)ctx->d_part_d.p;
    np.part_i = (long long*)ctx->d_part_i.p;
    np.tickets = (unsigned*)ctx->d_tickets.p;
    np.inst_ids = (const int32_t*)ctx->d_finst.p;
    np.seg_sizes = (const long long*)ctx->d_fsizes.p;
    np.seg = ctx->f_seg;
    np.grid_x = mps_nn_grid_x((long long)ctx->f_seg, ctx->num_sms);
    int cap_grid = ctx->num_sms * 16 / M;
    if (cap_grid < 1) cap_grid = 1;
    if (np.grid_x > cap_grid) np.grid_x = cap_grid;
    MPS_CUDA_CHECK(mps_launch_nn_scan(np, st), ctx);
    MPS_CUDA_CHECK(mps_launch_idx_to_parent(M, np.out_idx, (int32_t*)ctx->d_fparents.p, st), ctx);
    ctx->launches += 5;
    ctx->nn_ready = false;
    // 4. extend: K rollouts per instance, best by the active object's distance, appended to the tree
    std::vector<uint32_t> bf;
    if (flags) bf.resize((size_t)M);
    rc = forest_extend_launch(ctx, M, K, L, false, true, in->weights != nullptr, true, 0, in->n_goals, 1, best, nullptr,
                              nullptr, flags ? bf.data() : nullptr);
    if (rc) return rc;
    if (flags) memcpy(flags, bf.data(), sizeof(uint32_t) * (size_t)M);
    return MPS_OK;
}

int mps_forest_read(mps_ctx* ctx, int32_t inst, int64_t first, int64_t n, float* states, int32_t* parents,
                    float* controls)
{
    if (!ctx || inst < 0 || first < 0 || n < 0) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_forest_read");
    if (!ctx->forest.states || inst >= ctx->f_ninst) return mps_fail(ctx, MPS_ERR_ARG, "no such instance");
    if ((size_t)(first + n) > ctx->f_seg) return mps_fail(ctx, MPS_ERR_ARG, "range beyond the tree capacity");
    if (n == 0) return MPS_OK;
    const TreeView& t = ctx->forest;
    const size_t col = (size_t)inst * ctx->f_seg + (size_t)first;
    // dimension-major rows -> node-major host arrays: one strided 2-D copy per array
    if (states) {
        std::vector<float> tmp((size_t)n * 3 * t.B);
        MPS_CUDA_CHECK(cudaMemcpy2DAsync(tmp.data(), (size_t)n * sizeof(float), t.states + col, t.cap * sizeof(float),
                                         (size_t)n * sizeof(float), 3 * t.B, cudaMemcpyDeviceToHost, ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        for (int64_t i = 0; i < n; ++i)
            for (int r = 0; r < 3 * t.B; ++r) states[(size_t)i * 3 * t.B + r] = tmp[(size_t)r * n + i];
    }
    if (controls) {
        std::vector<float> tmp((size_t)n * MPS_CONTROL_DIM);
        MPS_CUDA_CHECK(cudaMemcpy2DAsync(tmp.data(), (size_t)n * sizeof(float), t.controls + col, t.cap * sizeof(float),
                                         (size_t)n * sizeof(float), MPS_CONTROL_DIM, cudaMemcpyDeviceToHost, ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        for (int64_t i = 0; i < n; ++i)
            for (int c = 0; c < MPS_CONTROL_DIM; ++c) controls[(size_t)i * MPS_CONTROL_DIM + c] = tmp[(size_t)c * n + i];
    }
    if (parents) {
        MPS_CUDA_CHECK(cudaMemcpyAsync(parents, t.parent + col, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    }
    return MPS_OK;
}

// ---- packed winner record -----------------------------------------------------------------------

int mps_extend_record_size(mps_ctx* ctx, int64_t* bytes)
{
    if (!ctx || !bytes) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "no extend uploaded");
    *bytes = (int64_t)record_bytes(ctx->scene.n_bodies, ctx->ext_L);
    return MPS_OK;
}

int mps_extend_record_to(mps_ctx* ctx, void* dst_device)
{
    if (!ctx || !dst_device) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "no extend uploaded");
    MPS_CUDA_CHECK(cudaMemcpyAsync(dst_device, ctx->d_record.p, record_bytes(ctx->scene.n_bodies, ctx->ext_L),
                                   cudaMemcpyDeviceToDevice, ctx->stream),
                   ctx);
    return MPS_OK;
}

int mps_extend_merge_records(mps_ctx* ctx, const void* gathered_device, int32_t world, int32_t has_dest,
                             int32_t compare_f32, int32_t append_to_tree, int32_t parent, mps_extend_best* best_out)
{
    if (!ctx || !gathered_device) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_extend_merge_records");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "no extend uploaded");
    if (world < 1 || world > 1024) return mps_fail(ctx, MPS_ERR_ARG, "world must be in [1, 1024]");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    const int B = ctx->scene.n_bodies, L = ctx->ext_L;
    MergeParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.gathered = (const unsigned char*)gathered_device;
    mp.world = world;
    mp.L = L;
    mp.B = B;
    mp.rec_bytes = record_bytes(B, L);
    mp.has_dest = has_dest ? 1 : 0;
    mp.compare_f32 = compare_f32 ? 1 : 0;
    mp.record = (unsigned char*)ctx->d_record.p;
    mp.head_out = nullptr;
    mp.append = append_to_tree ? 1 : 0;
    mp.parent = parent;
    if (mp.append) {
        int rc = tree_ensure(ctx, ctx->tree_upper + L);
        if (rc) return rc;
        mp.tree_size = (long long*)ctx->d_tree_size.p;
        mp.tree_cap = ctx->tree.cap;
        mp.tree_states = ctx->tree.states;
        mp.tree_controls = ctx->tree.controls;
        mp.tree_parent = ctx->tree.parent;
        mp.tree_slice = ctx->tree.slice;
        ctx->tree_upper += L;
    }
    MPS_CUDA_CHECK(mps_launch_merge_records(mp, ctx->stream), ctx);
    ctx->launches += 1;
    if (best_out) {
        // only the 24-byte head returns to the host
        int rc = ensure_stage(ctx, sizeof(mps_extend_best));
        if (rc) return rc;
        MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->h_stage, ctx->d_record.p, sizeof(mps_extend_best), cudaMemcpyDeviceToHost, ctx->stream), ctx);
        MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
        memcpy(best_out, ctx->h_stage, sizeof(mps_extend_best));
        if (mp.append) return tree_sync_size(ctx);
    }
    return MPS_OK;
}

int mps_extend_head(mps_ctx* ctx, mps_extend_best* best_out)
{
    if (!ctx || !best_out) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument to mps_extend_head");
    if (!ctx->extend_ready) return mps_fail(ctx, MPS_ERR_STATE, "no extend uploaded");
    int rc = ensure_stage(ctx, sizeof(mps_extend_best));
    if (rc) return rc;
    MPS_CUDA_CHECK(cudaMemcpyAsync(ctx->h_stage, ctx->d_record.p, sizeof(mps_extend_best), cudaMemcpyDeviceToHost, ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    memcpy(best_out, ctx->h_stage, sizeof(mps_extend_best));
    if (ctx->tree.states && ctx->tree_upper != ctx->tree_n) return tree_sync_size(ctx);
    return MPS_OK;
}

// ---- timing -----------------------------------------------------------------------------------

int mps_ctx_set_kernel_timing(mps_ctx* ctx, int32_t enabled)
{
    if (!ctx) return mps_fail(nullptr, MPS_ERR_ARG, "ctx is NULL");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    if (enabled && ctx->ev0.empty()) {
        ctx->ev0.resize(MPS_EVENT_RING);
        ctx->ev1.resize(MPS_EVENT_RING);
        for (int i = 0; i < MPS_EVENT_RING; ++i) {
            MPS_CUDA_CHECK(cudaEventCreate(&ctx->ev0[i]), ctx);
            MPS_CUDA_CHECK(cudaEventCreate(&ctx->ev1[i]), ctx);
        }
    }
    ctx->timing = enabled != 0;
    ctx->ev_count = 0;
    return MPS_OK;
}

int mps_extend_kernel_times(mps_ctx* ctx, float* ms, int32_t max_n, int32_t* n)
{
    if (!ctx || !ms || !n) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    int cnt = ctx->ev_count < max_n ? ctx->ev_count : max_n;
    for (int i = 0; i < cnt; ++i) MPS_CUDA_CHECK(cudaEventElapsedTime(&ms[i], ctx->ev0[i], ctx->ev1[i]), ctx);
    *n = cnt;
    ctx->ev_count = 0;
    return MPS_OK;
}

int mps_measure_fp32_peak(mps_ctx* ctx, float* tflops)
{
    if (!ctx || !tflops) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    MPS_CUDA_CHECK(cudaSetDevice(ctx->device), ctx);
    MPS_CUDA_CHECK(mps_run_fma_peak(ctx->num_sms, ctx->stream, tflops), ctx);
    return MPS_OK;
}

int mps_ctx_stream(mps_ctx* ctx, void** cuda_stream)
{
    if (!ctx || !cuda_stream) return mps_fail(ctx, MPS_ERR_ARG, "NULL argument");
    *cuda_stream = (void*)ctx->stream;
    return MPS_OK;
}


int mps_time_launches(mps_ctx* ctx, int32_t what, int32_t iters, float* elapsed_ms)
{
    if (!ctx || !elapsed_ms || iters < 1) return mps_fail(ctx, MPS_ERR_ARG, "bad argument to mps_time_launches");
    cudaEvent_t e0, e1;
    MPS_CUDA_CHECK(cudaEventCreate(&e0), ctx);
    MPS_CUDA_CHECK(cudaEventCreate(&e1), ctx);
    MPS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream), ctx);
    MPS_CUDA_CHECK(cudaEventRecord(e0, ctx->stream), ctx);
    int rc = MPS_OK;
    if (what == 2 && !ctx->topk_ready) rc = mps_fail(ctx, MPS_ERR_STATE, "no top-k query to repeat");
    for (int i = 0; i < iters && rc == MPS_OK; ++i) {
        if (what == 0) {
            rc = mps_extend_launch(ctx);
        } else if (what == 1) {
            rc = mps_tree_nearest_launch(ctx);
        } else {
            ctx->nn_chain = false;
            cudaError_t le = mps_launch_nn_topk(ctx->topk, ctx->stream);
            if (le != cudaSuccess) rc = mps_fail_cuda(ctx, le, "mps_launch_nn_topk", __LINE__);
            ctx->launches++;
        }
    }
    cudaError_t e = cudaEventRecord(e1, ctx->stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    if (e != cudaSuccess) return mps_fail_cuda(ctx, e, "event timing", __LINE__);
    *elapsed_ms = ms;
    return MPS_OK;
}

}  // extern "C"
