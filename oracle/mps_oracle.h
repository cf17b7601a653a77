/*
 * mps_oracle.h — CPU oracle of the kinodynamic-RRT extend step.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this.  The
 * product (manipulation_planning_suite_b200/, libmps_b200.so) never does.
 *
 * PARITY STATUS — read DESIGN.md §2:
 *   * control flow, ramp control, distance, goal, policy, bounds, best-of-K,
 *     chains, slices: restated from the reference sources cited per function
 *     and pinned by known-answer vectors derived from those in-tree formulas
 *     (tests/golden/).
 *   * the physics step itself (sim_env / Box2D in the reference, absent from
 *     /root/reference, un-vendored, un-pinned): "PARITY UNPINNED".  The
 *     oracle implements the spec "mps-planar-v3" written for this repo.
 */
#ifndef MPS_ORACLE_H
#define MPS_ORACLE_H

#include "../include/mps_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* util/Math.cpp:8-44 */
void orc_normalize_orientation_d(double* value);
void orc_normalize_orientation_f(float* value);
float orc_shortest_direction_so2(float val_1, float val_2);

/* RampVelocityControl.cpp:98-158 */
typedef struct orc_ramp {
    float v[3], a[3];
    float t_acc, t_plateau, t_rest;
} orc_ramp;
void orc_ramp_set(orc_ramp* r, const float params[5], const float accel_limits[3]);
void orc_ramp_velocity(const orc_ramp* r, float t, float vel[3]);
float orc_ramp_max_duration(const orc_ramp* r);
/* RampVelocityControl.cpp:68-79 */
void orc_ramp_from_local_twist(const float robot_pose[3], const float ltwist[4], float params[5]);

/* SimEnvWorldStateDistanceMeasure.cpp:34-51 over OMPL compound-space semantics */
double orc_distance(const mps_scene* sc, const float* a, const float* b, const float* weights, uint32_t mask);
/* ObjectsRelocationGoal.cpp:108-123 */
double orc_goal_distance(const float* state, const mps_goal* goals, int n_goals);
int orc_goal_satisfied(const float* state, const mps_goal* goals, int n_goals);
/* CollisionPolicy::collisionAllowed SimEnvState.cpp:1375-1394; b >= n_bodies means static b - n_bodies */
int orc_collision_allowed(const mps_scene* sc, int a, int b);

/* deterministic sin/cos of the physics spec (not libm) */
void orc_sincos(float a, float* s, float* c);

/* SimEnvValidityChecker::isValid SimEnvState.cpp:1441-1502; returns 1 if valid, *flags says why not */
int orc_is_valid(const mps_scene* sc, const float* state, uint32_t* flags);

/* SimEnvStatePropagator::propagate SimEnvStatePropagator.cpp:42-118.
 * control[4] (rest time) is written back.  in/out may alias.  Returns the bool. */
int orc_propagate(const mps_scene* sc, const float* state_in, float* control, float* state_out,
                  uint32_t* flags, int32_t* n_steps);

/* K chains best-of (NaiveControlSampler.cpp:34-71, RRT.cpp:449-489,531-565). Same structs as the ABI. */
int orc_extend(const mps_scene* sc, const mps_extend_in* in, const mps_extend_out* out);

/* exact linear-scan nearest neighbour over node-major states [n][3B]; ties -> lowest index */
int64_t orc_nn_linear(const mps_scene* sc, const float* states, const int32_t* slice_ids, int64_t n,
                      const float* query, const float* weights, uint32_t mask, int32_t slice_filter,
                      double* best_dist);

/* per-step trace for debugging parity: after every physics step the B x (x,y,th,vx,vy,w) are appended */
int orc_propagate_trace(const mps_scene* sc, const float* state_in, float* control, float* trace,
                        int32_t max_steps, int32_t* n_steps);

/* physics-only entry: one step from explicit poses+velocities (unit tests of the spec) */
int orc_step_once(const mps_scene* sc, float* pose /*[3B]*/, float* vel /*[3B]*/, const float u[3],
                  int32_t* n_manifolds);

/* contact query at a state: writes up to max_m manifolds (a, b, count, nx, ny, sep0, sep1) */
int orc_contacts(const mps_scene* sc, const float* state, int32_t* ab, float* data, int32_t max_m);

/* mps_gnat.c: restatement of ompl::NearestNeighborsGNAT, the structure behind the reference's _tree
 * (pushing/algorithm/RRT.cpp:46-51,202): the CPU baseline of the nearest-neighbour queries */
typedef struct orc_gnat orc_gnat;
orc_gnat* orc_gnat_create(const mps_scene* sc, const float* weights, uint32_t mask);
void orc_gnat_free(orc_gnat* g);
int orc_gnat_add(orc_gnat* g, const float* state);
int orc_gnat_size(const orc_gnat* g);
int orc_gnat_nearest(orc_gnat* g, const float* query, double* dist_out);
long orc_gnat_distance_evaluations(const orc_gnat* g);

/* mps_oracle_mt.c: pthread fan-out used by the host-CPU baseline */
int orc_extend_mt(const mps_scene* sc, const mps_extend_in* in, const mps_extend_out* out, int n_threads);
int orc_nn_linear_batch_mt(const mps_scene* sc, const float* states, const int32_t* slice_ids, int64_t n,
                           const float* queries, int32_t Q, const float* weights, const uint32_t* masks,
                           const int32_t* filters, int64_t* indices, double* dists, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
