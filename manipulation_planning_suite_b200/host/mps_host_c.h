/* mps_host_c.h — C entry points of libmps_host.so (the C++ planner classes of mps_host.hpp driven
 * through one call), used by the Python tests / benchmarks and by multi-seed sweeps. */
#ifndef MPS_HOST_C_H
#define MPS_HOST_C_H

#include <stdint.h>

#include "../../include/mps_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mps_host_params {
    int32_t algorithm;           /* 0 Naive, 1 OracleRRT, 2 SliceOracleRRT, 3 HybridActionRRT (AlgorithmType) */
    float time_out;              /* CPU seconds, util/Time.cpp:19-24 */
    float goal_bias, target_bias, robot_bias;
    int32_t num_control_samples;
    float action_randomness, state_noise;
    int32_t do_slice_ball_projection;
    int32_t device;
    uint64_t seed;
    int32_t max_iterations;      /* 0 = unlimited */
    float velocity_limits[3];
    float duration_limits[2];
    int32_t shortcut_type;       /* PlanningProblem::ShortcutType: 0 Naive, 1 Oracle, 2 none, 3 Local, 4 LocalOracle */
    float max_shortcut_time;     /* CPU seconds */
    int32_t shortcut_batch_size; /* candidates per batched launch; 1 = the reference's one-at-a-time evaluation */
} mps_host_params;

typedef struct mps_host_result {
    int32_t solved, reproducible;
    int32_t num_iterations, num_state_propagations, num_samples, num_nearest_neighbor_queries;
    float runtime_cpu, runtime_wall;
    int32_t tree_size, num_slices, path_len;
    float cost_before_shortcut, cost_after_shortcut;   /* PlanningStatistics, Interfaces.h:46-49 */
    int32_t reproducible_after_shortcut;
    int32_t path_len_before_shortcut;
    int32_t shortcut_candidates, shortcut_launches, shortcut_propagations, shortcut_accepted;
} mps_host_result;

void mps_host_default_params(mps_host_params* p);
/* plans from `start` to the relocation goals; path_states [max_path][3B], path_controls [max_path][5] */
int mps_host_plan(const mps_scene* scene, const float* start, const mps_goal* goals, int32_t n_goals,
                  const mps_host_params* params, mps_host_result* result, float* path_states, float* path_controls,
                  int32_t max_path, char* err, int32_t err_len);
/* shortcuts a given solution path (OraclePushPlanner::solve's shortcut step, OraclePushPlanner.cpp:229-238):
 * in_states [n_path][3B] / in_controls [n_path][5] as written by mps_host_plan */
int mps_host_shortcut(const mps_scene* scene, const mps_goal* goals, int32_t n_goals, const mps_host_params* params,
                      const float* in_states, const float* in_controls, int32_t n_path, mps_host_result* result,
                      float* path_states, float* path_controls, int32_t max_path, char* err, int32_t err_len);
/* RampComputer::steer (oracle/RampComputer.cpp:40-92) */
int mps_host_ramp_steer(const float* vel_limits, const float* accel_limits, const float* duration_limits,
                        const float* current, const float* desired, float* controls_out, int32_t max_controls);
/* SliceBasedOracleRRT's slice-ball projection (README.md:33-34, do_slice_ball_projection; declared but not
 * implemented in the reference snapshot - this repo's definition): `sample` [3B] is moved onto the ball of `radius`
 * around `center` under the arrangement metric (every body but the robot) when it lies outside; returns 1 when it
 * was inside (unchanged), 0 when it was projected, < 0 on a bad argument */
int mps_host_project_slice_ball(const mps_scene* scene, const float* weights, int32_t robot_id, float radius,
                                const float* center, float* sample);
/* HumanOracle::getMaximalPushingDistance (oracle/HumanOracle.cpp:161-164) */
float mps_host_max_pushing_distance(void);

#ifdef __cplusplus
}
#endif
#endif
