// mps_host.hpp — host side of the drop-in: C++ classes shaped like the reference's own interfaces for
// the extend path, implemented on top of the C ABI (include/mps_b200.h).  Nothing here computes physics
// or distances on the CPU: propagation, validity and nearest-neighbour all go to the device through the
// ABI; what stays on the host is what the reference also keeps outside the hot loop (sampling, oracle
// control generation, planner bookkeeping).
//
// Reference interfaces mirrored (paths relative to the reference root, same names / argument meaning):
//   mps::planner::ompl::control::RampVelocityControl{,Space,Sampler}   ompl/control/RampVelocityControl.{h,cpp}
//   mps::planner::ompl::control::SimEnvStatePropagator::propagate      ompl/control/SimEnvStatePropagator.cpp:42-118
//   mps::planner::ompl::control::NaiveControlSampler::sampleTo         ompl/control/NaiveControlSampler.cpp:34-71
//   mps::planner::ompl::state::SimEnvValidityChecker::isValid          ompl/state/SimEnvState.cpp:1441-1462
//   mps::planner::ompl::state::SimEnvWorldStateDistanceMeasure         ompl/state/SimEnvWorldStateDistanceMeasure.cpp
//   mps::planner::ompl::state::goal::ObjectsRelocationGoal             ompl/state/goal/ObjectsRelocationGoal.cpp
//   ompl::NearestNeighbors<MotionPtr> (add/nearest/clear/size)         as used in pushing/algorithm/RRT.cpp:46-51,202,239
//   mps::planner::pushing::oracle::{RampComputer,HumanOracle,OracleControlSampler}   pushing/oracle/*.cpp
//   mps::planner::pushing::algorithm::{RearrangementRRT, NaiveRearrangementRRT, HybridActionRRT,
//                                      OracleRearrangementRRT, SliceBasedOracleRRT}  pushing/algorithm/RRT.cpp:35-884
// Programming errors throw std::logic_error / std::runtime_error like the reference; ABI errors are
// turned into std::runtime_error carrying mps_last_error().
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <ctime>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mps_b200.h"

namespace mps {
namespace planner {

// ------------------------------------------------------------------------------------------- util
namespace util {
namespace math {
// util/Math.cpp:8-44
void normalize_orientation(double& value);
void normalize_orientation(float& value);
float shortest_direction_so2(float val_1, float val_2);
}  // namespace math

namespace random {
// The reference uses one process-global ompl::RNG (util/Random.cpp:9-16).  Here: one generator per
// planner instance (planner instances are independent), same distribution helpers.
class RNG {
public:
    explicit RNG(uint64_t seed = 1) : _gen(seed) {}
    double uniform01() { return (double)(_gen() >> 11) * (1.0 / 9007199254740992.0); }
    double uniformReal(double lo, double hi) { return lo + (hi - lo) * uniform01(); }
    int uniformInt(int lo, int hi) { return lo + (int)std::floor(uniform01() * (double)(hi - lo + 1)); }  // inclusive
    double gaussian01();
    double gaussian(double mean, double stddev) { return mean + stddev * gaussian01(); }
    void uniformInBall(double r, unsigned int dim, double* value);

private:
    std::mt19937_64 _gen;
};
typedef std::shared_ptr<RNG> RNGPtr;
}  // namespace random

namespace time {
// util/Time.cpp:15-33: the reference times out on process CPU time (std::clock)
class Timer {
public:
    void startTimer(float time_out);
    bool timeOutExceeded() const;
    float stopTimer();

private:
    std::clock_t _start = 0;
    float _time_out = 0.0f;
};
}  // namespace time
}  // namespace util

class DeviceContext;
typedef std::shared_ptr<DeviceContext> DeviceContextPtr;

// Owner of one mps_ctx (one planner instance on one device)
class DeviceContext {
public:
    DeviceContext(const mps_scene& scene, int device);
    ~DeviceContext();
    DeviceContext(const DeviceContext&) = delete;
    DeviceContext& operator=(const DeviceContext&) = delete;
    mps_ctx* get() const { return _ctx; }
    const mps_scene& scene() const { return _scene; }
    int device() const { return _device; }
    void check(int rc, const char* where) const;

private:
    mps_ctx* _ctx = nullptr;
    mps_scene _scene;
    int _device = 0;
};

namespace ompl {
// ------------------------------------------------------------------------------------------ state
namespace state {
// SimEnvWorldState in the semi-dynamic (position only) setting: (x, y, theta) per object
struct SimEnvWorldState {
    std::vector<float> values;  // 3 * n_objects
    explicit SimEnvWorldState(unsigned int n_objects = 0) : values(3 * (size_t)n_objects, 0.0f) {}
    unsigned int getNumObjects() const { return (unsigned int)(values.size() / 3); }
    float* getObjectState(unsigned int i) { return values.data() + 3 * (size_t)i; }
    const float* getObjectState(unsigned int i) const { return values.data() + 3 * (size_t)i; }
};

class SimEnvWorldStateSpace {
public:
    explicit SimEnvWorldStateSpace(const mps_scene& scene) : _scene(scene) {}
    unsigned int getNumObjects() const { return (unsigned int)_scene.n_bodies; }
    SimEnvWorldState* allocState() const { return new SimEnvWorldState(getNumObjects()); }
    void freeState(SimEnvWorldState* s) const { delete s; }
    void copyState(SimEnvWorldState* dst, const SimEnvWorldState* src) const { dst->values = src->values; }
    // SimEnvWorldStateSpace::copySubState SimEnvState.cpp:1081-1092
    void copySubState(SimEnvWorldState* s1, const SimEnvWorldState* s2, unsigned int obj_id) const;
    void sampleUniform(util::random::RNG& rng, SimEnvWorldState* s) const;
    void sampleUniformObject(util::random::RNG& rng, float* obj_state) const;
    const mps_scene& scene() const { return _scene; }

private:
    mps_scene _scene;
};
typedef std::shared_ptr<SimEnvWorldStateSpace> SimEnvWorldStateSpacePtr;

// SimEnvWorldStateDistanceMeasure: weights + active flags; evaluates the reference's double expression
class SimEnvWorldStateDistanceMeasure {
public:
    explicit SimEnvWorldStateDistanceMeasure(SimEnvWorldStateSpacePtr space, const std::vector<float>& weights = {});
    void setWeights(const std::vector<float>& weights);
    double distance(const SimEnvWorldState* a, const SimEnvWorldState* b) const;
    void setActive(unsigned int i, bool active) { _active.at(i) = active; }
    void setAll(bool active) { std::fill(_active.begin(), _active.end(), active); }
    bool isActive(unsigned int i) const { return _active.at(i); }
    uint32_t mask() const;
    const std::vector<float>& weights() const { return _weights; }

private:
    SimEnvWorldStateSpacePtr _space;
    std::vector<float> _weights;
    std::vector<bool> _active;
};
typedef std::shared_ptr<SimEnvWorldStateDistanceMeasure> SimEnvWorldStateDistanceMeasurePtr;

// SimEnvValidityChecker: the collision policy is part of the scene the device context was created with
class SimEnvValidityChecker {
public:
    explicit SimEnvValidityChecker(DeviceContextPtr ctx) : _ctx(ctx) {}
    bool isValid(const SimEnvWorldState* state) const;
    // batched form: one kernel launch for n states (the "next" row A15 / §8f rank 2)
    void isValidBatch(const std::vector<SimEnvWorldState>& states, std::vector<uint8_t>& valid) const;

private:
    DeviceContextPtr _ctx;
};
typedef std::shared_ptr<SimEnvValidityChecker> SimEnvValidityCheckerPtr;

namespace goal {
struct RelocationGoalSpecification {
    unsigned int id = 0;  // object index in the state
    float goal_position[2] = {0.0f, 0.0f};
    float position_tolerance = 0.1f;
};

class ObjectsRelocationGoal {
public:
    ObjectsRelocationGoal(SimEnvWorldStateSpacePtr space, SimEnvValidityCheckerPtr validity,
                          const std::vector<RelocationGoalSpecification>& goals, util::random::RNGPtr rng);
    double distanceGoal(const SimEnvWorldState* st) const;            // ObjectsRelocationGoal.cpp:108-123
    bool isSatisfied(const SimEnvWorldState* st) const;               // distanceGoal <= epsilon
    void sampleGoal(SimEnvWorldState* state) const;                   // :67-102 (<= 100 attempts)
    unsigned int sampleTargetObjectIndex() const;                     // :133-136 (last target twice as likely)
    bool canSample() const { return true; }
    const std::vector<mps_goal>& abiGoals() const { return _abi_goals; }

private:
    SimEnvWorldStateSpacePtr _space;
    SimEnvValidityCheckerPtr _validity;
    std::vector<RelocationGoalSpecification> _goals;
    std::vector<mps_goal> _abi_goals;
    std::vector<unsigned int> _target_indices;
    util::random::RNGPtr _rng;
    unsigned int _max_num_attempts = 100;
};
typedef std::shared_ptr<ObjectsRelocationGoal> ObjectsRelocationGoalPtr;
}  // namespace goal
}  // namespace state

// ---------------------------------------------------------------------------------------- control
namespace control {
class RampVelocityControlSpace;
typedef std::shared_ptr<RampVelocityControlSpace> RampVelocityControlSpacePtr;

// RampVelocityControl for the planar robot: (vx, vy, omega) + plateau duration + rest time
class RampVelocityControl {
public:
    explicit RampVelocityControl(const float accel_limits[3]);
    void setMaxVelocities(const float vel[3], float duration);
    void setParameters(const float params[5]);
    void getParameters(float params[5]) const;
    void setFromLocalTwist(const float robot_pose[3], const float ltwist[4]);
    void getVelocity(float dt, float vel[3]) const;
    float getMaxDuration() const;
    float getAccelerationTime() const { return _acceleration_duration; }
    float getRestTime() const { return _rest_time; }
    void setRestTime(float t) { _rest_time = t; }
    void addRestTime(float dt) { _rest_time += dt; }

private:
    void computeRamp();
    float _accel_limits[3];
    float _max_velocities[3] = {0, 0, 0};
    float _accelerations[3] = {0, 0, 0};
    float _plateau_duration = 0.0f, _rest_time = 0.0f, _acceleration_duration = 0.0f;
};

class RampVelocityControlSpace {
public:
    struct ControlLimits {
        float velocity_limits[3] = {0.6f, 0.6f, 1.5f};
        float acceleration_limits[3] = {2.0f, 2.0f, 4.0f};
        float duration_limits[2] = {0.1f, 1.0f};  // OraclePushPlanner.cpp:70-71
    };
    // one joint sub-space sampled in a ball: indices (0, 1) with norm limit (OracleParsing.cpp:121-127)
    RampVelocityControlSpace(const ControlLimits& limits, bool xy_ball, float xy_norm_limit)
        : _limits(limits), _xy_ball(xy_ball), _xy_norm_limit(xy_norm_limit) {}
    const ControlLimits& limits() const { return _limits; }
    RampVelocityControl* allocControl() const { return new RampVelocityControl(_limits.acceleration_limits); }
    void freeControl(RampVelocityControl* c) const { delete c; }
    void nullControl(float params[5]) const;
    // RampVelocityControlSampler::sample RampVelocityControl.cpp:353-394
    void sample(util::random::RNG& rng, float params[5]) const;

private:
    ControlLimits _limits;
    bool _xy_ball;
    float _xy_norm_limit;
};

// SimEnvStatePropagator: bool propagate(const State*, Control* (rest time mutated), State*)
class SimEnvStatePropagator {
public:
    explicit SimEnvStatePropagator(DeviceContextPtr ctx) : _ctx(ctx) {}
    bool propagate(const state::SimEnvWorldState* state, float control[5], state::SimEnvWorldState* result) const;
    bool canPropagateBackward() const { return false; }
    DeviceContextPtr context() const { return _ctx; }

private:
    DeviceContextPtr _ctx;
};
typedef std::shared_ptr<SimEnvStatePropagator> SimEnvStatePropagatorPtr;

// NaiveControlSampler: best of k sampled controls; the k propagations are one kernel launch
class NaiveControlSampler {
public:
    NaiveControlSampler(DeviceContextPtr ctx, RampVelocityControlSpacePtr control_space,
                        state::SimEnvWorldStateDistanceMeasurePtr distance, util::random::RNGPtr rng, unsigned int k);
    // returns 1 when a valid control was found (control / dest overwritten with the best), else 0 and dest := source
    unsigned int sampleTo(float control[5], const state::SimEnvWorldState* source, state::SimEnvWorldState* dest);
    void setK(unsigned int k) { _k = k > 0 ? k : 1; }
    unsigned int getK() const { return _k; }

private:
    DeviceContextPtr _ctx;
    RampVelocityControlSpacePtr _control_space;
    state::SimEnvWorldStateDistanceMeasurePtr _distance;
    util::random::RNGPtr _rng;
    unsigned int _k;
    std::vector<float> _controls;
};
}  // namespace control

// --------------------------------------------------------------------------------------- planning
namespace planning {
namespace essentials {
// Motion = tree node: state + control + parent (Essentials.cpp:20-93); `index` is its column in the device tree
struct Motion {
    state::SimEnvWorldState state;
    std::array<float, 5> control{{0, 0, 0, 0, 0}};
    std::shared_ptr<Motion> parent;
    int64_t index = -1;
    explicit Motion(unsigned int n_objects) : state(n_objects) {}
};
typedef std::shared_ptr<Motion> MotionPtr;

struct Path {
    std::vector<MotionPtr> motions;
    void initBacktrackMotion(MotionPtr last);  // Essentials.cpp:98-194
    unsigned int getNumMotions() const { return (unsigned int)motions.size(); }
};
typedef std::shared_ptr<Path> PathPtr;
}  // namespace essentials
}  // namespace planning
}  // namespace ompl

// Device-resident nearest-neighbour structure with the NearestNeighbors<MotionPtr> surface the
// reference uses (setDistanceFunction is replaced by an explicit (weights, mask) pair per query because
// the metric runs on the device).
class DeviceNearestNeighbors {
public:
    explicit DeviceNearestNeighbors(DeviceContextPtr ctx) : _ctx(ctx) {}
    void clear();
    std::size_t size() const { return _motions.size(); }
    void add(const ompl::planning::essentials::MotionPtr& m, int32_t slice_id = -1);
    // registers nodes the device appended itself (fused extend): host bookkeeping only
    void adopt(const ompl::planning::essentials::MotionPtr& m, int64_t index);
    ompl::planning::essentials::MotionPtr nearest(const ompl::state::SimEnvWorldState& query,
                                                  const std::vector<float>& weights, uint32_t mask,
                                                  int32_t slice_filter = -1, double* distance = nullptr) const;
    void list(std::vector<ompl::planning::essentials::MotionPtr>& out) const { out = _motions; }
    DeviceContextPtr context() const { return _ctx; }

private:
    DeviceContextPtr _ctx;
    std::vector<ompl::planning::essentials::MotionPtr> _motions;
};

// ----------------------------------------------------------------------------------------- oracle
namespace pushing {
namespace oracle {
struct ObjectData {  // oracle/Oracle.h:16-22
    float mass = 0, inertia = 0, mu = 0, width = 0, height = 0;
};

// RampComputer::steer (oracle/RampComputer.cpp:40-92): straight-line ramp decomposition
class RampComputer {
public:
    explicit RampComputer(ompl::control::RampVelocityControlSpacePtr control_space) : _control_space(control_space) {}
    void steer(const float current_robot_state[3], const float desired_robot_state[3],
               std::vector<std::array<float, 5>>& control_params) const;

private:
    ompl::control::RampVelocityControlSpacePtr _control_space;
};
typedef std::shared_ptr<RampComputer> RobotOraclePtr;

// HumanOracle (oracle/HumanOracle.cpp:37-172): closed-form push heuristic
struct HumanOracleParameters {  // HumanOracle::Parameters, oracle/HumanOracle.cpp:10-18
    float pushability_covariance[3] = {0.1f, 0.1f, 0.78f};  // diagonal
    float optimal_push_distance = 0.2f;
    float push_distance_tolerance = 0.04f;
    float push_angle_tolerance = 0.03f;
};

class HumanOracle {
public:
    typedef HumanOracleParameters Parameters;
    HumanOracle(RobotOraclePtr robot_oracle, const std::vector<ObjectData>& object_data, util::random::RNGPtr rng,
                const Parameters& params = Parameters());
    float predictPushability(const float cur_obj[3], const float next_obj[3]) const;
    float predictFeasibility(const float cur_robot[3], const float cur_obj[3], const float next_obj[3]) const;
    void predictAction(const float cur_robot[3], const float cur_obj[3], const float next_obj[3], unsigned int obj_id,
                       float control[5]) const;
    void sampleFeasibleState(const float cur_obj[3], const float next_obj[3], unsigned int obj_id, float new_robot[3]);
    float getMaximalPushingDistance() const;
    static void relativeSE2(const float state[3], const float ref[3], float out[3]);

private:
    Parameters _params;
    RobotOraclePtr _robot_steerer;
    std::vector<ObjectData> _object_data;
    util::random::RNGPtr _rng;
};
typedef std::shared_ptr<HumanOracle> PushingOraclePtr;

// OracleControlSampler (oracle/OracleControlSampler.cpp:52-281)
class OracleControlSampler {
public:
    OracleControlSampler(ompl::state::SimEnvWorldStateSpacePtr space, ompl::control::RampVelocityControlSpacePtr cs,
                         PushingOraclePtr pushing_oracle, RobotOraclePtr robot_oracle, unsigned int robot_id,
                         util::random::RNGPtr rng);
    void sampleFeasibleState(ompl::state::SimEnvWorldState* x_state, const ompl::state::SimEnvWorldState* x_prime_state,
                             unsigned int local_target_obj, float p_uniform = 0.0f);
    bool steerRobot(std::vector<std::array<float, 5>>& controls, const ompl::state::SimEnvWorldState* source,
                    const ompl::state::SimEnvWorldState* dest);
    bool steerPush(std::vector<std::array<float, 5>>& controls, const ompl::state::SimEnvWorldState* source,
                   const ompl::state::SimEnvWorldState* dest, unsigned int obj_id, float action_noise = 0.0f);
    void randomControl(std::vector<std::array<float, 5>>& controls);

private:
    ompl::state::SimEnvWorldStateSpacePtr _space;
    ompl::control::RampVelocityControlSpacePtr _control_space;
    PushingOraclePtr _pushing_oracle;
    RobotOraclePtr _robot_oracle;
    unsigned int _robot_id;
    util::random::RNGPtr _rng;
};
typedef std::shared_ptr<OracleControlSampler> OracleControlSamplerPtr;
}  // namespace oracle

// -------------------------------------------------------------------------------------- algorithm
namespace algorithm {
struct PlanningStatistics {  // algorithm/Interfaces.h:39-127
    unsigned int num_iterations = 0, num_state_propagations = 0, num_samples = 0, num_nearest_neighbor_queries = 0;
    float runtime = 0.0f;       // CPU seconds like the reference
    float wall_time = 0.0f;     // additionally: wall clock seconds
    bool success = false;
    bool reproducible = false;
    float cost_before_shortcut = 0.0f, cost_after_shortcut = 0.0f;  // Interfaces.h:46-49
    bool reproducible_after_shortcut = false;
};

// ompl::base::ParamSet stand-in: string-typed set / get with declared setters (RRT.cpp:145-169)
class ParamSet {
public:
    void declareParam(const std::string& name, std::function<void(const std::string&)> setter,
                      std::function<std::string()> getter);
    bool setParam(const std::string& name, const std::string& value);
    bool getParam(const std::string& name, std::string& value) const;
    bool hasParam(const std::string& name) const { return _setters.count(name) > 0; }

private:
    std::map<std::string, std::function<void(const std::string&)>> _setters;
    std::map<std::string, std::function<std::string()>> _getters;
};

struct PlanningQuery {  // algorithm/Interfaces.h:131-150
    ompl::state::goal::ObjectsRelocationGoalPtr goal_region;
    ompl::state::SimEnvWorldState start_state;
    unsigned int robot_id = 0;
    float time_out = 60.0f;
    std::function<bool()> stopping_condition = []() { return false; };
    ompl::planning::essentials::PathPtr path;
    std::shared_ptr<ParamSet> parameters = std::make_shared<ParamSet>();
    std::vector<float> weights;
    unsigned int max_iterations = 0;  // 0 = unlimited (extension: deterministic stopping for tests)
};
typedef std::shared_ptr<PlanningQuery> PlanningQueryPtr;

struct SpaceInformation {  // what the reference passes around as ompl::control::SpaceInformationPtr
    DeviceContextPtr ctx;
    ompl::state::SimEnvWorldStateSpacePtr state_space;
    ompl::control::RampVelocityControlSpacePtr control_space;
    ompl::state::SimEnvValidityCheckerPtr validity_checker;
    ompl::control::SimEnvStatePropagatorPtr state_propagator;
    util::random::RNGPtr rng;
};
typedef std::shared_ptr<SpaceInformation> SpaceInformationPtr;

class RearrangementRRT {
protected:
    struct PlanningBlackboard {
        PlanningQueryPtr pq;
        PlanningStatistics stats;
        unsigned int robot_id = 0;
        explicit PlanningBlackboard(PlanningQueryPtr pq_) : pq(pq_) {}
    };
    typedef ompl::planning::essentials::MotionPtr MotionPtr;

public:
    explicit RearrangementRRT(SpaceInformationPtr si);
    virtual ~RearrangementRRT() = default;
    bool plan(PlanningQueryPtr pq);
    bool plan(PlanningQueryPtr pq, PlanningStatistics& stats);  // RRT.cpp:79-143
    virtual PlanningQueryPtr createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                                 const ompl::state::SimEnvWorldState& start_state,
                                                 unsigned int robot_id, float timeout = 60.0f);
    virtual bool extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int active_obj_id,
                        MotionPtr& last_motion, PlanningBlackboard& pb) = 0;
    virtual bool sample(MotionPtr motion, unsigned int& target_obj_id, PlanningBlackboard& pb);  // :171-190
    virtual void selectTreeNode(const MotionPtr& sample, MotionPtr& selected_node, unsigned int& active_obj_id,
                                bool sample_is_goal, PlanningBlackboard& pb);                 // :192-204
    float getGoalBias() const { return _goal_bias; }
    float getTargetBias() const { return _target_bias; }
    float getRobotBias() const { return _robot_bias; }
    void setGoalBias(float v) { _goal_bias = v; }
    void setTargetBias(float v) { _target_bias = v; }
    void setRobotBias(float v) { _robot_bias = v; }
    std::shared_ptr<util::time::Timer> timer_ptr = std::make_shared<util::time::Timer>();
    std::size_t treeSize() const { return _tree->size(); }

protected:
    virtual void setup(PlanningQueryPtr pq, PlanningBlackboard& blackboard);                  // :60-71
    virtual void addToTree(MotionPtr new_motion, MotionPtr parent, PlanningBlackboard& pb);   // :234-244
    unsigned int sampleActiveObject(const PlanningBlackboard& pb) const;                      // :246-257
    void sampleValidState(ompl::state::SimEnvWorldState* s);  // allocValidStateSampler stand-in (:53)
    MotionPtr getNewMotion() { return std::make_shared<ompl::planning::essentials::Motion>(_state_space->getNumObjects()); }
    uint32_t fullMask() const;

    SpaceInformationPtr _si;
    ompl::state::SimEnvWorldStateSpacePtr _state_space;
    ompl::state::SimEnvWorldStateDistanceMeasurePtr _distance_measure;
    util::random::RNGPtr _rng;
    std::shared_ptr<DeviceNearestNeighbors> _tree;
    float _goal_bias = 0.2f, _target_bias = 0.25f, _robot_bias = 0.0f;
};

class NaiveRearrangementRRT : public RearrangementRRT {
public:
    NaiveRearrangementRRT(SpaceInformationPtr si, unsigned int k = 10);
    PlanningQueryPtr createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                         const ompl::state::SimEnvWorldState& start_state, unsigned int robot_id,
                                         float timeout = 60.0f) override;
    bool extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int active_obj_id, MotionPtr& last_motion,
                PlanningBlackboard& pb) override;  // RRT.cpp:358-390
    unsigned int getNumControlSamples() const { return _num_control_samples; }
    void setNumControlSamples(unsigned int n) { _num_control_samples = n; }

protected:
    void setup(PlanningQueryPtr pq, PlanningBlackboard& blackboard) override;
    unsigned int _num_control_samples;
    std::vector<float> _controls;
};

class HybridActionRRT : public RearrangementRRT {
public:
    HybridActionRRT(SpaceInformationPtr si, oracle::PushingOraclePtr pushing_oracle, oracle::RobotOraclePtr robot_oracle,
                    unsigned int robot_id);
    PlanningQueryPtr createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                         const ompl::state::SimEnvWorldState& start_state, unsigned int robot_id,
                                         float timeout = 60.0f) override;
    bool extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int active_obj_id, MotionPtr& last_motion,
                PlanningBlackboard& pb) override;  // RRT.cpp:449-489
    unsigned int getNumControlSamples() const { return _num_control_samples; }
    void setNumControlSamples(unsigned int k) { _num_control_samples = k; }
    float getActionRandomness() const { return _action_randomness; }
    void setActionRandomness(float ar) { _action_randomness = ar; }

protected:
    void sampleActionSequence(std::vector<std::array<float, 5>>& controls, MotionPtr start,
                              ompl::state::SimEnvWorldState* dest, PlanningBlackboard& pb);  // :491-529
    oracle::OracleControlSamplerPtr _oracle_sampler;
    unsigned int _num_control_samples = 10;
    float _action_randomness = 0.0001f;
    static const int kMaxChain = MPS_MAX_CHAIN;
};

class OracleRearrangementRRT : public RearrangementRRT {
public:
    OracleRearrangementRRT(SpaceInformationPtr si, oracle::PushingOraclePtr pushing_oracle,
                           oracle::RobotOraclePtr robot_oracle, unsigned int robot_id);
    PlanningQueryPtr createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                         const ompl::state::SimEnvWorldState& start_state, unsigned int robot_id,
                                         float timeout = 60.0f) override;
    void selectTreeNode(const MotionPtr& sample, MotionPtr& selected_node, unsigned int& active_obj_id,
                        bool sample_is_goal, PlanningBlackboard& pb) override;  // RRT.cpp:632-661
    bool extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int active_obj_id, MotionPtr& last_motion,
                PlanningBlackboard& pb) override;  // :663-696
    float getActionRandomness() const { return _action_randomness; }
    void setActionRandomness(float ar) { _action_randomness = ar; }
    float getStateNoise() const { return _feasible_state_noise; }
    void setStateNoise(float sn) { _feasible_state_noise = sn; }

protected:
    void extendStep(const std::vector<std::array<float, 5>>& controls, const MotionPtr& start_motion,
                    MotionPtr& result_motion, PlanningBlackboard& pb, bool& extension_success, bool& goal_reached);  // :698-735
    oracle::OracleControlSamplerPtr _oracle_sampler;
    float _action_randomness = 0.0001f;
    float _feasible_state_noise = 0.00001f;
};

// A slice of the joint configuration space: all nodes with (almost) the same object arrangement
struct Slice {  // algorithm/Interfaces.h:181-193
    ompl::planning::essentials::MotionPtr repr;
    std::vector<ompl::planning::essentials::MotionPtr> slice_samples_list;
    int32_t id = -1;
};
typedef std::shared_ptr<Slice> SlicePtr;

class SliceBasedOracleRRT : public OracleRearrangementRRT {
public:
    SliceBasedOracleRRT(SpaceInformationPtr si, oracle::PushingOraclePtr pushing_oracle,
                        oracle::RobotOraclePtr robot_oracle, unsigned int robot_id);
    void selectTreeNode(const MotionPtr& sample, MotionPtr& selected_node, unsigned int& active_obj_id,
                        bool sample_is_goal, PlanningBlackboard& pb) override;  // RRT.cpp:783-827
    std::size_t numSlices() const { return _slices_list.size(); }
    // slice-ball projection of a sampled arrangement onto the ball of radius max_pushing_distance around the
    // nearest slice (README.md:33-34; declared but not implemented in the reference snapshot: our definition)
    void setSliceBallProjection(bool on, float radius) { _do_slice_ball_projection = on; _slice_ball_radius = radius; }
    bool projectSliceBall(const ompl::state::SimEnvWorldState& center, ompl::state::SimEnvWorldState* sample) const;
    // the projection itself (needs no device): true = the sample was inside the ball and is unchanged
    static bool projectSliceBall(const ompl::state::SimEnvWorldStateDistanceMeasure& measure, unsigned int n_objects,
                                 unsigned int robot_id, float radius, const ompl::state::SimEnvWorldState& center,
                                 ompl::state::SimEnvWorldState* sample);

protected:
    void setup(PlanningQueryPtr pq, PlanningBlackboard& pb) override;                   // :766-781
    void addToTree(MotionPtr new_motion, MotionPtr parent, PlanningBlackboard& pb) override;  // :829-844
    SlicePtr getSlice(const ompl::state::SimEnvWorldState& state, PlanningBlackboard& pb, double* distance) const;
    uint32_t arrangementMask(unsigned int robot_id) const;
    std::shared_ptr<DeviceNearestNeighbors> _slices_nn;  // one node per slice representative, own device context
    std::vector<SlicePtr> _slices_list;
    oracle::PushingOraclePtr _pushing_oracle;
    double _min_slice_distance = 0.0;
    bool _do_slice_ball_projection = false;
    float _slice_ball_radius = 0.0f;
    unsigned int _robot_id_cached = 0;
};
// ----------------------------------------------------------------------------------- shortcutting
// CostFunction (ompl/planning/Essentials.cpp:197-211) and ActionDurationCost (pushing/Costs.cpp:12-17)
class CostFunction {
public:
    virtual ~CostFunction() = default;
    // cost of taking the action stored in `second` from the state in `first`
    virtual double cost(const ompl::planning::essentials::Motion& first,
                        const ompl::planning::essentials::Motion& second) const = 0;
    // accumulates over the motions 0..limit (all when limit < 0); starts with prev = motion 0, so the control
    // stored in motion 0 is counted as well (Essentials.cpp:203-209)
    virtual double cost(const ompl::planning::essentials::Path& path, int limit = -1) const;
};
typedef std::shared_ptr<CostFunction> CostFunctionPtr;

class ActionDurationCost : public CostFunction {
public:
    explicit ActionDurationCost(const float accel_limits[3]);
    using CostFunction::cost;
    double cost(const ompl::planning::essentials::Motion& first,
                const ompl::planning::essentials::Motion& second) const override;
    double duration(const std::array<float, 5>& control) const;  // RampVelocityControl::getMaxDuration

private:
    float _accel_limits[3];
};

// Shortcutter (RRT.h:313-412, RRT.cpp:889-1034).  The reference tries one candidate (start_id, end_id) at a
// time: build the replacement actions on the host, forward propagate them and the rest of the old path, accept
// when the goal is still reached at lower cost.  Here every candidate of a round is an independent control
// chain with its own start state, so a whole round is propagated by batched launches (mps_extend with
// `starts`, `stop_at_goal`) and the candidates are then examined in the reference's order; the first
// acceptable one is taken, exactly the candidate the sequential loop would have accepted.
class Shortcutter {
public:
    struct ShortcutQuery {  // RRT.h:319-335
        ompl::state::goal::ObjectsRelocationGoalPtr goal_region;
        CostFunctionPtr cost_function;
        unsigned int robot_id = 0;
        std::function<bool()> stopping_condition = []() { return false; };
        float cost_before_shortcut = 0.0f, cost_after_shortcut = 0.0f;
        ShortcutQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region_in, CostFunctionPtr cost_function_in,
                      unsigned int robot_id_in)
            : goal_region(goal_region_in), cost_function(cost_function_in), robot_id(robot_id_in)
        {
        }
    };
    struct Statistics {
        unsigned int num_candidates = 0;      // candidate shortcuts evaluated
        unsigned int num_launches = 0;        // batched propagate launches
        unsigned int num_propagations = 0;    // controls propagated on the device
        unsigned int num_accepted = 0;
    };
    explicit Shortcutter(SpaceInformationPtr si);
    virtual ~Shortcutter() = default;
    virtual void shortcut(ompl::planning::essentials::PathPtr iopath, ShortcutQuery& sq, float max_time) = 0;
    virtual std::string getName() const = 0;
    const Statistics& statistics() const { return _stats; }
    // candidates per batched launch (upper bound; 1 reproduces the reference's one-at-a-time evaluation)
    void setBatchSize(unsigned int n) { _batch_size = n > 0 ? n : 1; }

protected:
    typedef ompl::planning::essentials::Motion Motion;
    typedef ompl::planning::essentials::MotionPtr MotionPtr;
    typedef ompl::planning::essentials::Path Path;
    typedef ompl::planning::essentials::PathPtr PathPtr;
    typedef std::array<float, 5> ControlParams;
    // one candidate shortcut being propagated: `appended` grows by the motions that propagated successfully
    struct Candidate {
        unsigned int start_id = 0, end_id = 0, object_id = 0;
        int stage = 0;                        // next stage nextControls() is asked for
        std::deque<ControlParams> pending;    // controls of the current stage not yet propagated
        std::vector<MotionPtr> appended;      // forwardPropagatePath's path->append(new_motion)
        ompl::state::SimEnvWorldState last;   // state the next control starts from
        bool valid = true, goal = false, done = false;
    };
    // fills c.pending with the controls of stage c.stage given c.last; returns false when there is no further
    // stage (c stays valid), may set c.valid = false to abort the candidate (e.g. oracle gave no control)
    virtual bool nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq) = 0;
    // forwardPropagatePath (RRT.cpp:946-1034) for all candidates at once
    void propagateCandidates(std::vector<Candidate>& cands, const Path& current_path, ShortcutQuery& sq);
    void initCandidate(Candidate& c, const Path& current_path, unsigned int start_id, unsigned int end_id) const;
    void finish(PathPtr iopath, const Path& current_path, ShortcutQuery& sq, double initial_cost, double cost) const;

    SpaceInformationPtr _si;
    util::time::Timer _timer;
    Statistics _stats;
    unsigned int _batch_size = 256;
};
typedef std::shared_ptr<Shortcutter> ShortcutterPtr;

class NaiveShortcutter : public Shortcutter {  // RRT.cpp:1054-1171
public:
    NaiveShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle);
    void shortcut(PathPtr iopath, ShortcutQuery& sq, float max_time) override;
    std::string getName() const override { return "NaiveShortcutter"; }

protected:
    bool nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq) override;
    void createAllPairs(std::vector<std::pair<unsigned int, unsigned int>>& all_pairs, unsigned int n) const;
    oracle::RobotOraclePtr _robot_oracle;
};

class LocalShortcutter : public Shortcutter {  // RRT.cpp:1177-1305
public:
    LocalShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle, unsigned int robot_id);
    void shortcut(PathPtr iopath, ShortcutQuery& sq, float max_time) override;
    std::string getName() const override { return "LocalShortcutter"; }

protected:
    bool nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq) override;
    oracle::RobotOraclePtr _robot_oracle;
    unsigned int _robot_id;
};

class LocalOracleShortcutter : public LocalShortcutter {  // RRT.cpp:1310-1411
public:
    LocalOracleShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle,
                           oracle::PushingOraclePtr pushing_oracle, unsigned int robot_id);
    std::string getName() const override { return "LocalOracleShortcutter"; }

protected:
    bool nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq) override;
    unsigned int selectObject(const ompl::state::SimEnvWorldState& start_state,
                              const ompl::state::SimEnvWorldState& end_state) const;  // :1326-1345
    oracle::OracleControlSamplerPtr _oracle_sampler;
};

class OracleShortcutter : public Shortcutter {  // RRT.cpp:1416-1688
public:
    OracleShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle, oracle::PushingOraclePtr pushing_oracle,
                      unsigned int robot_id);
    void shortcut(PathPtr iopath, ShortcutQuery& sq, float max_time) override;
    std::string getName() const override { return "OracleShortcutter"; }

protected:
    typedef std::pair<unsigned int, unsigned int> IndexPair;
    typedef std::deque<IndexPair> PairQueue;
    typedef std::vector<unsigned int> ObjectIds;
    typedef std::map<IndexPair, ObjectIds> PairMap;
    bool nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq) override;
    void fillPairQueue(PairQueue& all_pairs, unsigned int n) const;                                  // :1505-1519
    unsigned int selectObject(const Path& current_path, const IndexPair& current_pair, PairMap& pair_to_objects,
                              PairQueue& pair_queue, unsigned int robot_id) const;                   // :1521-1571
    oracle::OracleControlSamplerPtr _oracle_sampler;
};
}  // namespace algorithm

// ---------------------------------------------------------------------------------- OraclePushPlanner
enum class AlgorithmType { Naive = 0, OracleRRT = 1, SliceOracleRRT = 2, HybridActionRRT = 3 };
// PlanningProblem::ShortcutType, OraclePushPlanner.h:48-54 (same values)
enum class ShortcutType { NaiveShortcut = 0, OracleShortcut = 1, NoShortcut = 2, LocalShortcut = 3, LocalOracleShortcut = 4 };

struct PlanningProblem {  // OraclePushPlanner.cpp:40-91 (the fields the path reads)
    mps_scene scene;
    ompl::state::SimEnvWorldState start_state;
    unsigned int robot_id = 0;
    std::vector<ompl::state::goal::RelocationGoalSpecification> relocation_goals;
    ompl::control::RampVelocityControlSpace::ControlLimits control_limits;
    float planning_time_out = 60.0f;
    float goal_bias = 0.1f, robot_bias = 0.1f, target_bias = 0.1f;
    float action_noise = 0.01f, state_noise = 0.001f;
    bool do_slice_ball_projection = true;
    unsigned int num_control_samples = 10;
    AlgorithmType algorithm_type = AlgorithmType::Naive;
    ShortcutType shortcut_type = ShortcutType::LocalShortcut;  // OraclePushPlanner.cpp:87-88
    float max_shortcut_time = 5.0f;
    unsigned int shortcut_batch_size = 256;
    std::vector<float> object_weights;  // empty = 1.0 each (prepareDistanceWeights :610-622)
    uint64_t seed = 1;
    int device = 0;
    unsigned int max_iterations = 0;
    std::function<bool()> stopping_condition = []() { return false; };
};

struct PlanningSolution {
    ompl::planning::essentials::PathPtr path;
    bool solved = false;
    algorithm::PlanningStatistics stats;
};

class OraclePushPlanner {
public:
    bool setup(PlanningProblem& problem);       // OraclePushPlanner.cpp:107-173
    bool solve(PlanningSolution& solution);     // :175-243
    void shortcutSolution(PlanningSolution& solution);  // :229-238, also usable on a loaded solution
    bool verifySolution(PlanningSolution& solution);  // :495-519: replay every control, check the goal
    algorithm::SpaceInformationPtr getSpaceInformation() const { return _si; }
    std::shared_ptr<algorithm::RearrangementRRT> getAlgorithm() const { return _algorithm; }
    algorithm::ShortcutterPtr getShortcutter() const { return _shortcutter; }
    unsigned int pathLengthBeforeShortcut() const { return _path_len_before_shortcut; }
    ompl::state::goal::ObjectsRelocationGoalPtr getGoal() const { return _goal; }

private:
    PlanningProblem _problem;
    algorithm::SpaceInformationPtr _si;
    std::shared_ptr<algorithm::RearrangementRRT> _algorithm;
    algorithm::ShortcutterPtr _shortcutter;
    unsigned int _path_len_before_shortcut = 0;
    ompl::state::goal::ObjectsRelocationGoalPtr _goal;
    bool _is_initialized = false;
};
}  // namespace pushing
}  // namespace planner
}  // namespace mps
