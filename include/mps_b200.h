/*
 * mps_b200.h — C ABI of the B200-native kinodynamic-RRT extend step.
 *
 * This is the drop-in boundary for the one hot path of
 * JoshuaHaustein/manipulation_planning_suite: K push rollouts
 * + fused validity / goal / slice-ball tests + argmin, and the weighted
 * SE(2)^n nearest-neighbour scan over a device-resident tree.
 *
 * Plain C: pointers and sizes only, no C++ / torch types.  Every entry point
 * returns an int status (MPS_OK = 0, < 0 = error; text via mps_last_error()).
 * Rollout validity / goal flags travel in outputs, never in the status.
 * There is NO CPU fallback: without a CUDA device every call that needs one
 * fails with MPS_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   mps_extend / mps_propagate      <- SimEnvStatePropagator::propagate
 *                                      src/mps/planner/ompl/control/SimEnvStatePropagator.cpp:42-118
 *                                      + NaiveControlSampler::sampleTo  NaiveControlSampler.cpp:34-71
 *                                      + HybridActionRRT::extend loop   pushing/algorithm/RRT.cpp:449-489
 *   mps_is_valid_batch              <- SimEnvValidityChecker::isValid   ompl/state/SimEnvState.cpp:1441-1462
 *   mps_tree_*                      <- ompl::NearestNeighbors<MotionPtr> (add/nearest/clear/size/list)
 *                                      as used at pushing/algorithm/RRT.cpp:46-51,202,239 and
 *                                      SliceBasedOracleRRT RRT.cpp:766-884
 *   mps_scene (policy, bounds, …)   <- PlanningProblem defaults         pushing/OraclePushPlanner.cpp:40-91
 *   mps_distance                    <- SimEnvWorldStateDistanceMeasure::distance
 *                                      ompl/state/SimEnvWorldStateDistanceMeasure.cpp:34-51
 */
#ifndef MPS_B200_H
#define MPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPS_ABI_VERSION 2

#define MPS_MAX_BODIES 32   /* B: robot + movable objects (one warp lane per body) */
#define MPS_MAX_STATICS 16  /* S: static obstacles (walls, fixed boxes) */
#define MPS_MAX_GOALS 8
#define MPS_CONTROL_DIM 5   /* (vx, vy, omega, t_plateau, t_rest) RampVelocityControl.cpp:34-61 */
#define MPS_MAX_MANIFOLDS 64
#define MPS_MAX_CHAIN 64     /* controls per chain (L) */

/* status codes */
#define MPS_OK 0
#define MPS_ERR_ARG (-1)
#define MPS_ERR_NO_DEVICE (-2)
#define MPS_ERR_CUDA (-3)
#define MPS_ERR_CAPACITY (-4)
#define MPS_ERR_STATE (-5)

/* shapes */
#define MPS_SHAPE_BOX 0
#define MPS_SHAPE_DISC 1

/* per-propagation flags (one uint32 per propagated control) */
#define MPS_F_RAN (1u << 0)         /* this control was propagated at all            */
#define MPS_F_OK (1u << 1)          /* propagate() returned true                     */
#define MPS_F_GOAL (1u << 2)        /* goal_region.isSatisfied(result)               */
#define MPS_F_IN_BALL (1u << 3)     /* result inside the slice ball (if requested)   */
#define MPS_F_NOT_AT_REST (1u << 4) /* t_max exceeded before the world came to rest  */
#define MPS_F_REST_CONTACT (1u << 5)/* disallowed contact seen in the rest phase     */
#define MPS_F_BOUNDS (1u << 6)      /* result violates the workspace bounds          */
#define MPS_F_INFEASIBLE (1u << 7)  /* result penetrates deeper than pen_max         */
#define MPS_F_BAD_CONTACT (1u << 8) /* result has a contact the policy forbids       */
#define MPS_F_OVERFLOW (1u << 9)    /* more touching pairs than max_manifolds        */
#define MPS_F_ACTIVE_CONTACT (1u << 10) /* strict_active_contacts only               */
#define MPS_F_JAMMED (1u << 11)     /* rest phase aborted: penetration > 1.25 * slop for jam_steps steps */

/*
 * Scene = everything that is fixed for a planning problem: bodies, statics,
 * collision policy, workspace bounds, control acceleration limits, distance
 * sub-weights and the constants of the planar push physics ("mps-planar-v3",
 * DESIGN.md §3).  Body order is the order of the reference's
 * SimEnvWorldStateSpace sub-spaces (SimEnvState.cpp:1027-1050).
 */
typedef struct mps_scene {
    int32_t n_bodies;  /* B, 2..MPS_MAX_BODIES */
    int32_t robot_id;  /* index of the velocity-controlled body */
    int32_t n_statics; /* S, 0..MPS_MAX_STATICS */
    int32_t vel_iters; /* velocity iterations per step */

    /* movable bodies (robot is kinematic: its mass / inertia are ignored) */
    int32_t shape[MPS_MAX_BODIES];
    float hx[MPS_MAX_BODIES]; /* box half extent x, or disc radius */
    float hy[MPS_MAX_BODIES]; /* box half extent y (unused for discs) */
    float mass[MPS_MAX_BODIES];
    float inertia[MPS_MAX_BODIES];
    float mu_ground[MPS_MAX_BODIES];  /* support-plane friction */
    float fr_radius[MPS_MAX_BODIES];  /* lever arm of the angular support friction */
    float mu_contact[MPS_MAX_BODIES]; /* body-body Coulomb friction, mixed sqrt(mu_a mu_b) */

    /* static obstacles */
    int32_t s_shape[MPS_MAX_STATICS];
    float s_x[MPS_MAX_STATICS];
    float s_y[MPS_MAX_STATICS];
    float s_theta[MPS_MAX_STATICS];
    float s_hx[MPS_MAX_STATICS];
    float s_hy[MPS_MAX_STATICS];
    float s_mu_contact[MPS_MAX_STATICS];

    /* workspace bounds on (x, y) of every body: constructLimits SimEnvState.cpp:1225-1287 */
    float x_min, x_max, y_min, y_max;

    /* collision policy resolved to ids: CollisionPolicy SimEnvState.cpp:1365-1394 */
    uint32_t pair_forbidden[MPS_MAX_BODIES]; /* bit j of row i (symmetric) */
    uint32_t static_forbidden;               /* bit i: body i must not touch statics */

    /* RampVelocityControlSpace acceleration limits (vx, vy, omega) */
    float accel_limits[3];
    float _pad0;

    /* DistanceWeights SimEnvState.h:224-231 (double there, double here) */
    double position_weight;    /* 1.0  */
    double orientation_weight; /* 0.08 */

    /* physics constants */
    float dt;           /* physics time step */
    float t_max;        /* semi-dynamic rest budget, SimEnvStatePropagator.h:31 */
    float gravity;      /* for support friction caps */
    float slop;         /* allowed penetration */
    float baumgarte;    /* position-error feedback factor */
    float max_bias_vel; /* cap on the push-out velocity */
    float pen_max;      /* penetration beyond which a state is physically infeasible */
    float v_rest;       /* atRest linear threshold */
    float w_rest;       /* atRest angular threshold */
    int32_t max_manifolds;          /* touching-pair capacity per rollout, <= MPS_MAX_MANIFOLDS */
    int32_t strict_active_contacts; /* 0 = replicate SimEnvStatePropagator.cpp:79 shadowing quirk */
    int32_t jam_steps;              /* rest phase gives up after this many consecutive steps deeper than 1.25 * slop; 0 = never */
} mps_scene;

/* RelocationGoalSpecification reduced to what distanceGoal reads (ObjectsRelocationGoal.cpp:108-123) */
typedef struct mps_goal {
    int32_t body;
    float x, y, tol;
} mps_goal;

/* Inputs of one extend: K rollouts (chains of <= L controls) from one start state. */
typedef struct mps_extend_in {
    const float* start;    /* [3B] (x, y, theta) per body */
    const float* controls; /* [K][L][5] */
    const int32_t* n_ctrl; /* [K] controls used per chain, or NULL = L for all */
    int32_t K, L;
    const float* dest;     /* [3B] distance target, or NULL (then no argmin) */
    const float* weights;  /* [B] per-object weights, or NULL = 1 */
    uint32_t mask;         /* active-object bit mask of the distance */
    int32_t compare_f32;   /* 1: compare (float)distance like RRT.cpp:466-467; 0: double like NaiveControlSampler.cpp:53-55 */
    const mps_goal* goals; /* [n_goals] or NULL */
    int32_t n_goals;
    int32_t k_offset;      /* global index of rollout 0 (multi-GPU sharding) */
    const float* ball_center; /* [3B] slice-ball centre or NULL */
    float ball_radius;
    int32_t append_to_tree; /* 1: the winner chain is appended to the ctx tree on device */
    int32_t parent;         /* tree index of `start` (used when append_to_tree) */
    int32_t stop_at_goal;   /* 1 = a chain ends at its first goal state (the shortcutters' forwardPropagatePath,
                               RRT.cpp:966-968,1011-1013); 0 = the whole chain is propagated */
    const float* starts;    /* [K][3B] optional per-chain start states (candidate replays of the shortcutters,
                               RRT.cpp:946-1034); NULL = every chain starts from `start` */
} mps_extend_in;

/* Winner of an extend (strict <, first index wins; NaiveControlSampler.cpp:55, RRT.cpp:467). */
typedef struct mps_extend_best {
    int32_t k;         /* global index of the winning rollout, -1 if none valid */
    int32_t n_ok;      /* successful controls in its chain */
    int32_t goal_step; /* first chain step whose result satisfies the goal, -1 if none */
    int32_t tree_index;/* index of the first appended node, -1 if not appended */
    double distance;   /* d(result, dest) of the winner */
} mps_extend_best;

/* Optional outputs; any pointer may be NULL. */
typedef struct mps_extend_out {
    mps_extend_best* best;
    float* best_states;    /* [L][3B] states along the winning chain */
    float* best_controls;  /* [L][5] controls of the winning chain, t_rest filled in */
    uint32_t* best_flags;  /* [L] */
    /* full dumps (parity tests, data generation) */
    float* all_states;     /* [K][L][3B] */
    float* all_rest;       /* [K][L] rest times written back into the controls */
    uint32_t* all_flags;   /* [K][L] */
    double* all_dist;      /* [K] d(last ok state, dest); +inf when the chain has no ok step */
    int32_t* all_steps;    /* [K][L] physics steps executed (roofline accounting) */
} mps_extend_out;

typedef struct mps_ctx mps_ctx;

/* ---- library ---------------------------------------------------------- */
int mps_abi_version(void);
const char* mps_last_error(const mps_ctx* ctx); /* ctx may be NULL: last error of create */
int mps_device_count(int* count);

/* fills every field with the defaults of DESIGN.md §3 / SURVEY.md Appendix A */
void mps_scene_defaults(mps_scene* scene);

/* ---- context: one per (device, planner instance); owns all device memory */
int mps_ctx_create(const mps_scene* scene, int device, mps_ctx** out);
/* same, but work is enqueued on a caller-owned cudaStream_t (e.g. torch's current stream) */
int mps_ctx_create_on_stream(const mps_scene* scene, int device, void* cuda_stream, mps_ctx** out);
void mps_ctx_destroy(mps_ctx* ctx);
int mps_ctx_synchronize(mps_ctx* ctx);
int mps_ctx_scene(const mps_ctx* ctx, mps_scene* out);

/* ---- propagation ------------------------------------------------------ */
/* Whole extend with host buffers: H2D, kernels, D2H, synchronous. */
int mps_extend(mps_ctx* ctx, const mps_extend_in* in, const mps_extend_out* out);
/* K=1, L=1 form of the reference call: bool propagate(in, control, out); control[4] (rest time)
 * is mutated.  state_in and state_out may alias (OraclePushPlanner.cpp:510-512). *ok receives the bool. */
int mps_propagate(mps_ctx* ctx, const float* state_in, float* control, float* state_out, int32_t* ok,
                  uint32_t* flags);
/* Staged form for resident pipelines: upload once, launch many, fetch when needed. */
int mps_extend_upload(mps_ctx* ctx, const mps_extend_in* in);
int mps_extend_launch(mps_ctx* ctx);                       /* asynchronous on the ctx stream */
int mps_extend_fetch(mps_ctx* ctx, const mps_extend_out* out); /* synchronises */
/* ---- action sequences generated on the device ---------------------------------------------------------------
 * HybridActionRRT::sampleActionSequence (pushing/algorithm/RRT.cpp:491-529) for all K chains of an extend in one
 * kernel: per chain a die (random control with probability action_randomness), the object to move
 * (sampleActiveObject RRT.cpp:246-257, or fixed_object for OracleRRT), then either the transit primitive -
 * RampComputer::steer of the robot towards its sampled pose (oracle/RampComputer.cpp:40-92) - or the transfer
 * primitive - HumanOracle::sampleFeasibleState (oracle/HumanOracle.cpp:136-159), the steer to that pushing pose and
 * HumanOracle::predictAction from it (:111-134).  Random numbers are a hash of (seed, iteration, k_offset + chain,
 * purpose, index), so an extend uploads this struct instead of K x L x 20 bytes of controls, and a chain's actions
 * do not depend on how the chains are split over GPUs. */
typedef struct mps_hybrid_gen {
    uint64_t seed;
    int32_t iteration;
    int32_t k_offset;          /* global index of chain 0 (multi-GPU sharding; normally = mps_extend_in.k_offset) */
    int32_t fixed_object;      /* >= 0: every chain acts on this body; -1: sampled per chain */
    float action_randomness;   /* RRT.cpp:500 */
    float target_bias, robot_bias;
    int32_t n_targets;
    int32_t targets[MPS_MAX_GOALS];
    float velocity_limits[3];  /* per component (vx, vy, omega); the random control draws (vx, vy) from the disc of
                                  radius velocity_limits[0] and omega from +-velocity_limits[2] */
    float duration_min, duration_max;
    float optimal_push_distance, push_distance_tolerance, push_angle_tolerance; /* HumanOracle.cpp:10-18 */
    float object_size[MPS_MAX_BODIES]; /* max(width, height) of every body (HumanOracle.cpp:121) */
} mps_hybrid_gen;
/* mps_extend_upload with the controls and chain lengths produced on the device: in->controls and in->n_ctrl are
 * ignored, in->K chains of at most in->L controls are generated (a sequence that needs more is cut at L and
 * counted).  Then mps_extend_launch / mps_extend_fetch as usual. */
int mps_extend_upload_generated(mps_ctx* ctx, const mps_extend_in* in, const mps_hybrid_gen* gen);
/* the generated controls [K][L][5] and chain lengths [K] of the last mps_extend_upload_generated (for replay /
 * inspection; either may be NULL), and how many chains were cut at L */
int mps_extend_generated_controls(mps_ctx* ctx, float* controls, int32_t* n_ctrl, int32_t* n_truncated);
/* n controls of the device's RampVelocityControlSampler::sample (ompl/control/RampVelocityControl.cpp:353-394), the
 * sampler mps_forest_naive_step and the generator above draw random controls with: hash index gid, draws 0 .. n-1 */
int mps_sample_controls(mps_ctx* ctx, uint64_t seed, int32_t iteration, int64_t gid, int32_t n, float velocity_limit,
                        float omega_limit, float duration_min, float duration_max, float* controls);
/* counters of the last launch: [0] physics steps, [1] candidate pairs, [2] touching manifolds,
 * [3] solver colour passes, [4] kernel launches since ctx creation */
int mps_extend_counters(mps_ctx* ctx, uint64_t counters[8]);
/* SM clock cycles the last launch spent on each of its K chains (profiling aid: which chain is the tail) */
int mps_extend_chain_cycles(mps_ctx* ctx, int64_t* cycles);

/* ---- validity --------------------------------------------------------- */
/* valid[i] = isValid(states[i]); flags[i] (optional) says why not */
int mps_is_valid_batch(mps_ctx* ctx, const float* states, int32_t n, uint8_t* valid, uint32_t* flags);

/* ---- distance (host helper, double arithmetic of the reference) ------- */
int mps_distance(const mps_scene* scene, const float* a, const float* b, const float* weights,
                 uint32_t mask, double* out);
int mps_goal_distance(const mps_scene* scene, const float* state, const mps_goal* goals, int32_t n_goals,
                      double* out);

/* ---- tree: device-resident SoA + exact weighted NN -------------------- */
int mps_tree_reserve(mps_ctx* ctx, int64_t capacity);
int mps_tree_clear(mps_ctx* ctx);
int mps_tree_size(mps_ctx* ctx, int64_t* n);
/* append one node / many nodes; slice_id may be -1; control may be NULL (null control) */
int mps_tree_add(mps_ctx* ctx, const float* state, int32_t parent, const float* control, int32_t slice_id,
                 int64_t* index);
int mps_tree_add_batch(mps_ctx* ctx, const float* states, const int32_t* parents, const float* controls,
                       const int32_t* slice_ids, int64_t n, int64_t* first_index);
/* nearest node under d = sum_i mask_i w_i (pw |dxy| + ow arc(dtheta)), double arithmetic,
 * ties -> lowest index.  slice_filter >= 0 restricts to nodes of that slice; -1 = all.
 * One launch, no stream synchronise: the scan's CTAs post their partial results into pinned host memory and the
 * call folds them (same minimum as the device-side reduction of the staged / batched calls). */
int mps_tree_nearest(mps_ctx* ctx, const float* query, const float* weights, uint32_t mask,
                     int32_t slice_filter, int64_t* index, double* distance);
/* ompl::NearestNeighbors<MotionPtr>::nearestK of the interface the planners hold their tree through
 * (include/mps/planner/pushing/algorithm/RRT.h:149): the k <= 32 nearest nodes, ascending by (double distance,
 * node index) - the order a sequential double scan with strict < would produce.  *n_found <= k (fewer nodes in the
 * tree / the slice).  One streaming pass with warp-level top-k lists; exact (see csrc/nn_scan.cu). */
int mps_tree_nearest_k(mps_ctx* ctx, const float* query, const float* weights, uint32_t mask, int32_t slice_filter,
                       int32_t k, int64_t* indices, double* distances, int32_t* n_found);
/* n calls of mps_tree_nearest in a row (queries [n][3B]), timed by the host clock around the loop: the end-to-end
 * latency of the C entry point itself (host query in, index and distance out), without a language binding on top */
int mps_tree_nearest_timed(mps_ctx* ctx, const float* queries, int32_t n, const float* weights, uint32_t mask,
                           int32_t slice_filter, int64_t* indices, double* distances, double* seconds);
/* how many mps_tree_nearest_k calls had to be redone by exact successor scans (diagnostic) */
int mps_tree_nearest_k_fallbacks(mps_ctx* ctx, int64_t* count);
/* SliceBasedOracleRRT's two-level search (RRT.cpp:783-827: _slices_nn->nearest, then the slice's own
 * slice_samples_nn->nearest) as two scans chained on the device: the nearest slice representative in `reps` (a
 * second context on the same device whose node index is the slice id) under rep_mask, then the nearest node of
 * this context's tree under member_mask among the nodes of that slice.  No host round trip between the two; valid
 * when the sample is not adjusted between the levels (active object = robot).  index = -1 when the slice is empty. */
int mps_tree_nearest_two_level(mps_ctx* ctx, mps_ctx* reps, const float* query, const float* weights, uint32_t rep_mask,
                               uint32_t member_mask, int64_t* rep_index, double* rep_distance, int64_t* index,
                               double* distance);
/* Q queries, one launch per query batch; masks/filters per query may be NULL (= all / -1) */
int mps_tree_nearest_batch(mps_ctx* ctx, const float* queries, int32_t Q, const float* weights,
                           const uint32_t* masks, const int32_t* slice_filters, int64_t* indices,
                           double* distances);
/* staged form (queries resident in HBM) */
int mps_tree_nearest_upload(mps_ctx* ctx, const float* queries, int32_t Q, const float* weights,
                            const uint32_t* masks, const int32_t* slice_filters);
int mps_tree_nearest_launch(mps_ctx* ctx);
int mps_tree_nearest_fetch(mps_ctx* ctx, int64_t* indices, double* distances);
int mps_tree_get(mps_ctx* ctx, int64_t index, float* state, int32_t* parent, float* control,
                 int32_t* slice_id);
/* path root..index: indices[<=max_len], returns length in *len */
int mps_tree_backtrack(mps_ctx* ctx, int64_t index, int64_t* indices, int64_t max_len, int64_t* len);
/* fill the tree with n seeded uniform states directly on device (benchmarks; no H2D of 132 MB) */
int mps_tree_fill_random(mps_ctx* ctx, int64_t n, uint64_t seed, int32_t n_slices);
/* copy node states [first, first+n) back to the host, node-major [n][3B] */
int mps_tree_read_states(mps_ctx* ctx, int64_t first, int64_t n, float* states);

/* ---- forest: many independent planner instances in one context ----------- */
/* BASELINE config 5 (multi-seed / multi-problem sweeps): n_instances trees of equal capacity share one
 * dimension-major SoA; one launch serves a batch of M instances (SURVEY.md §8e "independent planner
 * instances").  Instance indices are local to the context; node indices are local to their tree. */
typedef struct mps_forest_extend_in {
    int32_t M, K, L;          /* M instances in this batch, K chains of <= L controls each */
    const int32_t* inst_ids;  /* [M] which tree */
    const int32_t* parents;   /* [M] node of that tree the chains start from (read on device) */
    const float* controls;    /* [M][K][L][5] */
    const int32_t* n_ctrl;    /* [M][K] or NULL */
    const float* dests;       /* [M][3B] or NULL */
    const uint32_t* masks;    /* [M] or NULL (all bodies) */
    const float* weights;     /* [B] or NULL */
    int32_t compare_f32;
    int32_t n_goals;
    const mps_goal* goals;    /* shared by all instances */
    int32_t append_to_tree;   /* winners appended to their own trees */
    int32_t _pad;
} mps_forest_extend_in;

int mps_forest_reserve(mps_ctx* ctx, int32_t n_instances, int32_t cap_per_instance);
/* every tree := { root }, roots [n_instances][3B] */
int mps_forest_reset(mps_ctx* ctx, const float* roots);
int mps_forest_sizes(mps_ctx* ctx, int64_t* sizes);
int mps_forest_nearest(mps_ctx* ctx, int32_t M, const int32_t* inst_ids, const float* queries, const float* weights,
                       const uint32_t* masks, int64_t* indices, double* distances);
/* best [M], best_states [M][L][3B], best_controls [M][L][5], best_flags [M][L]; any may be NULL */
int mps_forest_extend(mps_ctx* ctx, const mps_forest_extend_in* in, mps_extend_best* best, float* best_states,
                      float* best_controls, uint32_t* best_flags);
/* One whole NaiveRRT iteration for M instances without host random numbers or per-iteration uploads: on the
 * device, every instance draws its sample (goal-biased or uniform, the first valid of C candidate states:
 * RearrangementRRT::sample RRT.cpp:171-190, ObjectsRelocationGoal::sampleGoal :67-102), its active object
 * (sampleActiveObject RRT.cpp:246-257), finds the nearest node of its own tree, draws K controls
 * (RampVelocityControlSampler::sample RampVelocityControl.cpp:353-394), propagates them, appends the best.
 * Random numbers are a hash of (seed, iteration, instance_offset + instance id, purpose, index): they do not
 * depend on the other instances nor on how the instances are split over devices. */
typedef struct mps_forest_naive_in {
    int32_t M;                 /* active instances of this iteration */
    int32_t iteration;
    const int32_t* inst_ids;   /* [M] which tree */
    int64_t instance_offset;   /* global instance id = instance_offset + inst_ids[m] */
    uint64_t seed;
    int32_t K;                 /* num_control_samples */
    int32_t C;                 /* candidate states per sample */
    float goal_bias, target_bias, robot_bias;
    int32_t n_goals;
    const mps_goal* goals;     /* relocation goals shared by all instances */
    const float* weights;      /* [B] or NULL */
    float velocity_limit, omega_limit, duration_min, duration_max;  /* control limits */
} mps_forest_naive_in;
/* best [M] (k, n_ok, goal_step, tree_index, distance), flags [M] (MPS_F_* of the appended state); may be NULL */
int mps_forest_naive_step(mps_ctx* ctx, const mps_forest_naive_in* in, mps_extend_best* best, uint32_t* flags);
/* nodes [first, first + n) of one tree: states [n][3B], parents [n], controls [n][5]; any may be NULL */
int mps_forest_read(mps_ctx* ctx, int32_t inst, int64_t first, int64_t n, float* states, int32_t* parents,
                    float* controls);

/* ---- packed winner record (multi-GPU gather) ---------------------------- */
/* Layout: mps_extend_best | float best_states[L][3B] | float best_controls[L][5] | uint32 best_flags[L],
 * padded to 8 bytes.  mps_extend_record_to copies the record of the last launch device-to-device
 * (asynchronously on the ctx stream) into caller-owned device memory, e.g. the input of an NCCL
 * all-gather. */
int mps_extend_record_size(mps_ctx* ctx, int64_t* bytes);
int mps_extend_record_to(mps_ctx* ctx, void* dst_device);
/* Exchange step of the K-sharded extend, on the device (replaces the tail of the reference's sample loop,
 * NaiveControlSampler.cpp:53-70 / RRT.cpp:466-488, for chains that were propagated on other GPUs).
 * `gathered_device` = `world` packed records in rank order (the all-gather of every rank's record).  Enqueued on
 * the ctx stream: the global winner is picked by the reference's rule (strict <; equal keys -> lowest global
 * chain index), its record becomes this ctx's record (mps_extend_fetch / mps_extend_record_to see it), and when
 * append_to_tree is set - on the GPU that owns the tree - its chain is appended below `parent` (up to the first
 * goal state).  best_out != NULL: synchronises and returns the 24-byte head; NULL: asynchronous, read the head
 * later with mps_extend_head. */
int mps_extend_merge_records(mps_ctx* ctx, const void* gathered_device, int32_t world, int32_t has_dest,
                             int32_t compare_f32, int32_t append_to_tree, int32_t parent, mps_extend_best* best_out);
/* the 24-byte head of the ctx's current winner record (synchronises); nothing else crosses PCIe */
int mps_extend_head(mps_ctx* ctx, mps_extend_best* best_out);

/* ---- timing helper ---------------------------------------------------- */
/* Runs `iters` back-to-back launches of the uploaded extend (what=0), NN batch (what=1) or last top-k query (what=2)
 * on the ctx stream between two CUDA events and returns the elapsed milliseconds. */
int mps_time_launches(mps_ctx* ctx, int32_t what, int32_t iters, float* elapsed_ms);
/* Per-launch device time of the rollout kernel: when enabled, every mps_extend_launch brackets the
 * rollout kernel with CUDA events on the ctx stream (ring of 512 launches).  mps_extend_kernel_times
 * synchronises, writes the elapsed milliseconds of the recorded launches (oldest first) and clears the ring. */
int mps_ctx_set_kernel_timing(mps_ctx* ctx, int32_t enabled);
int mps_extend_kernel_times(mps_ctx* ctx, float* ms, int32_t max_n, int32_t* n);
/* FP32 FMA microbenchmark on the ctx device: the measured peak the rollout roofline is quoted against */
int mps_measure_fp32_peak(mps_ctx* ctx, float* tflops);
/* raw cudaStream_t the ctx enqueues on (for event timing by the caller) */
int mps_ctx_stream(mps_ctx* ctx, void** cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MPS_B200_H */
