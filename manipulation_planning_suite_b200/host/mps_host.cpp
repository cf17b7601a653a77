// mps_host.cpp — see mps_host.hpp.  Every propagation / validity / nearest-neighbour call below goes
// through the C ABI to the device; there is no CPU implementation of those in this file.
#include "mps_host.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <ctime>
#include <cstdio>

#include "mps_host_c.h"

namespace mps {
namespace planner {

// float -> string for the string-typed ParamSet (no iostreams: the library may be loaded into hosts
// whose C++ runtime state we do not own)
static std::string f2s(float v)
{
    char buf[64];
    snprintf(buf, sizeof(buf), "%.9g", (double)v);
    return std::string(buf);
}

static const double kPi = 3.14159265358979323846;
static const float kPiF = 3.14159265358979323846f;

// =========================================================================================== util
namespace util {
namespace math {
void normalize_orientation(double& value)
{
    double v = std::fmod(value, 2.0 * kPi);
    if (v < -kPi)
        v += 2.0 * kPi;
    else if (v >= kPi)
        v -= 2.0 * kPi;
    value = v;
}

void normalize_orientation(float& value)
{
    float v = std::fmod(value, 2.0f * kPiF);
    if (v < -kPiF)
        v += 2.0f * kPiF;
    else if (v >= kPiF)
        v -= 2.0f * kPiF;
    value = v;
}

float shortest_direction_so2(float val_1, float val_2)
{
    float value = val_2 - val_1;
    if (std::abs(value) > kPiF) {
        if (value > 0.0f)
            value -= 2.0f * kPiF;
        else
            value += 2.0f * kPiF;
    }
    return value;
}
}  // namespace math

namespace random {
double RNG::gaussian01()
{
    // Box-Muller on our own uniforms (std::normal_distribution is implementation defined)
    double u1 = uniform01(), u2 = uniform01();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * kPi * u2);
}

// ompl::RNG::uniformInBall: normalised gaussian direction, radius r * u^(1/dim)
void RNG::uniformInBall(double r, unsigned int dim, double* value)
{
    double norm = 0.0;
    for (unsigned int i = 0; i < dim; ++i) {
        value[i] = gaussian01();
        norm += value[i] * value[i];
    }
    norm = std::sqrt(norm);
    if (norm == 0.0) {
        value[0] = 1.0;
        norm = 1.0;
    }
    double radius = r * std::pow(uniform01(), 1.0 / (double)dim);
    for (unsigned int i = 0; i < dim; ++i) value[i] = radius * value[i] / norm;
}
}  // namespace random

namespace time {
void Timer::startTimer(float time_out)
{
    _start = std::clock();
    _time_out = time_out;
}
bool Timer::timeOutExceeded() const { return (float)(std::clock() - _start) / (float)CLOCKS_PER_SEC > _time_out; }
float Timer::stopTimer() { return (float)(std::clock() - _start) / (float)CLOCKS_PER_SEC; }
}  // namespace time
}  // namespace util

// ================================================================================== DeviceContext
DeviceContext::DeviceContext(const mps_scene& scene, int device) : _scene(scene), _device(device)
{
    int rc = mps_ctx_create(&scene, device, &_ctx);
    if (rc != MPS_OK) {
        const char* msg = mps_last_error(nullptr);
        throw std::runtime_error(std::string("[mps::planner::DeviceContext] ") + (msg ? msg : "mps_ctx_create failed"));
    }
}

DeviceContext::~DeviceContext() { mps_ctx_destroy(_ctx); }

void DeviceContext::check(int rc, const char* where) const
{
    if (rc != MPS_OK) {
        const char* msg = mps_last_error(_ctx);
        throw std::runtime_error(std::string("[") + where + "] " + (msg ? msg : "ABI call failed"));
    }
}

namespace ompl {
// ========================================================================================== state
namespace state {
void SimEnvWorldStateSpace::copySubState(SimEnvWorldState* s1, const SimEnvWorldState* s2, unsigned int obj_id) const
{
    for (int c = 0; c < 3; ++c) s1->values[3 * obj_id + c] = s2->values[3 * obj_id + c];
}

void SimEnvWorldStateSpace::sampleUniformObject(util::random::RNG& rng, float* o) const
{
    o[0] = (float)rng.uniformReal(_scene.x_min, _scene.x_max);
    o[1] = (float)rng.uniformReal(_scene.y_min, _scene.y_max);
    o[2] = (float)rng.uniformReal(-kPi, kPi);
}

void SimEnvWorldStateSpace::sampleUniform(util::random::RNG& rng, SimEnvWorldState* s) const
{
    for (unsigned int i = 0; i < getNumObjects(); ++i) sampleUniformObject(rng, s->getObjectState(i));
}

SimEnvWorldStateDistanceMeasure::SimEnvWorldStateDistanceMeasure(SimEnvWorldStateSpacePtr space,
                                                                 const std::vector<float>& weights)
    : _space(space), _weights(weights)
{
    _active = std::vector<bool>(space->getNumObjects(), true);
    if (_weights.empty()) _weights.resize(_active.size(), 1.0f);
}

void SimEnvWorldStateDistanceMeasure::setWeights(const std::vector<float>& weights)
{
    if (weights.size() == _weights.size()) _weights = weights;  // else: "Insufficient number of weights", kept as is
}

uint32_t SimEnvWorldStateDistanceMeasure::mask() const
{
    uint32_t m = 0;
    for (size_t i = 0; i < _active.size(); ++i)
        if (_active[i]) m |= 1u << i;
    return m;
}

double SimEnvWorldStateDistanceMeasure::distance(const SimEnvWorldState* a, const SimEnvWorldState* b) const
{
    double d = 0.0;
    int rc = mps_distance(&_space->scene(), a->values.data(), b->values.data(), _weights.data(), mask(), &d);
    if (rc != MPS_OK) throw std::logic_error("[mps::planner::ompl::state::SimEnvWorldStateDistanceMeasure::distance] bad state");
    return d;
}

bool SimEnvValidityChecker::isValid(const SimEnvWorldState* state) const
{
    uint8_t v = 0;
    _ctx->check(mps_is_valid_batch(_ctx->get(), state->values.data(), 1, &v, nullptr), "SimEnvValidityChecker::isValid");
    return v != 0;
}

void SimEnvValidityChecker::isValidBatch(const std::vector<SimEnvWorldState>& states, std::vector<uint8_t>& valid) const
{
    valid.assign(states.size(), 0);
    if (states.empty()) return;
    const size_t dim = states[0].values.size();
    std::vector<float> flat(states.size() * dim);
    for (size_t i = 0; i < states.size(); ++i) std::memcpy(flat.data() + i * dim, states[i].values.data(), dim * sizeof(float));
    _ctx->check(mps_is_valid_batch(_ctx->get(), flat.data(), (int32_t)states.size(), valid.data(), nullptr),
                "SimEnvValidityChecker::isValidBatch");
}

namespace goal {
ObjectsRelocationGoal::ObjectsRelocationGoal(SimEnvWorldStateSpacePtr space, SimEnvValidityCheckerPtr validity,
                                             const std::vector<RelocationGoalSpecification>& goals,
                                             util::random::RNGPtr rng)
    : _space(space), _validity(validity), _goals(goals), _rng(rng)
{
    for (auto& g : _goals) {
        if (g.id >= space->getNumObjects())
            throw std::runtime_error("[mps::planner::ompl::state::goal::ObjectRelocationGoal] Could not retrieve state space for target object.");
        mps_goal a;
        a.body = (int32_t)g.id;
        a.x = g.goal_position[0];
        a.y = g.goal_position[1];
        a.tol = g.position_tolerance;
        _abi_goals.push_back(a);
        _target_indices.push_back(g.id);
    }
}

double ObjectsRelocationGoal::distanceGoal(const SimEnvWorldState* st) const
{
    double d = 0.0;
    int rc = mps_goal_distance(&_space->scene(), st->values.data(), _abi_goals.data(), (int32_t)_abi_goals.size(), &d);
    if (rc != MPS_OK) throw std::logic_error("[mps::planner::ompl::state::goal::ObjectRelocationGoal::distanceGoal] bad state");
    return d;
}

bool ObjectsRelocationGoal::isSatisfied(const SimEnvWorldState* st) const
{
    return distanceGoal(st) <= std::numeric_limits<double>::epsilon();
}

void ObjectsRelocationGoal::sampleGoal(SimEnvWorldState* state) const
{
    // <= 100 attempts of "uniform state + targets in their goal discs", the first valid one wins.  The
    // candidates of a round are validated by ONE device launch (16 per round) instead of one world
    // save / set / collision query / restore each (SimEnvState.cpp:1441-1462).
    const unsigned int per_round = 16;
    unsigned int attempts = 0;
    std::vector<SimEnvWorldState> cands;
    std::vector<uint8_t> valid;
    SimEnvWorldState last(_space->getNumObjects());
    while (attempts < _max_num_attempts) {
        cands.clear();
        for (unsigned int i = 0; i < per_round && attempts + i < _max_num_attempts; ++i) {
            SimEnvWorldState s(_space->getNumObjects());
            _space->sampleUniform(*_rng, &s);
            for (const auto& g : _goals) {
                double values[2];
                _rng->uniformInBall(g.position_tolerance, 2, values);
                float* o = s.getObjectState(g.id);
                o[0] = g.goal_position[0] + (float)values[0];
                o[1] = g.goal_position[1] + (float)values[1];
            }
            cands.push_back(s);
        }
        attempts += (unsigned int)cands.size();
        _validity->isValidBatch(cands, valid);
        for (size_t i = 0; i < cands.size(); ++i) {
            if (valid[i]) {
                *state = cands[i];
                return;
            }
        }
        last = cands.back();
    }
    *state = last;  // "Failed to sample a valid goal, providing invalid one."
}

unsigned int ObjectsRelocationGoal::sampleTargetObjectIndex() const
{
    int idx = std::min(_rng->uniformInt(0, (int)_target_indices.size()), ((int)_target_indices.size()) - 1);
    return _target_indices.at((size_t)idx);
}
}  // namespace goal
}  // namespace state

// ======================================================================================== control
namespace control {
RampVelocityControl::RampVelocityControl(const float accel_limits[3])
{
    for (int i = 0; i < 3; ++i) _accel_limits[i] = accel_limits[i];
}

void RampVelocityControl::setMaxVelocities(const float vel[3], float duration)
{
    for (int i = 0; i < 3; ++i) _max_velocities[i] = vel[i];
    _plateau_duration = duration;
    computeRamp();
}

void RampVelocityControl::setParameters(const float p[5])
{
    setMaxVelocities(p, p[3]);
    setRestTime(p[4]);
}

void RampVelocityControl::getParameters(float p[5]) const
{
    for (int i = 0; i < 3; ++i) p[i] = _max_velocities[i];
    p[3] = _plateau_duration;
    p[4] = _rest_time;
}

void RampVelocityControl::setFromLocalTwist(const float robot_pose[3], const float ltwist[4])
{
    float heading = robot_pose[2] + ltwist[3];
    float p[5] = {ltwist[0] * std::cos(heading), ltwist[0] * std::sin(heading), ltwist[1], ltwist[2], 0.0f};
    setParameters(p);
}

void RampVelocityControl::getVelocity(float dt, float vel[3]) const
{
    if (dt < _acceleration_duration) {
        for (int i = 0; i < 3; ++i) vel[i] = dt * _accelerations[i];
    } else if (dt < _acceleration_duration + _plateau_duration) {
        for (int i = 0; i < 3; ++i) vel[i] = _max_velocities[i];
    } else if (dt < 2.0 * _acceleration_duration + _plateau_duration) {
        float s = dt - _acceleration_duration - _plateau_duration;
        for (int i = 0; i < 3; ++i) vel[i] = _max_velocities[i] - s * _accelerations[i];
    } else {
        for (int i = 0; i < 3; ++i) vel[i] = 0.0f;
    }
}

float RampVelocityControl::getMaxDuration() const { return 2.0f * _acceleration_duration + _plateau_duration + _rest_time; }

void RampVelocityControl::computeRamp()
{
    float max_accel_time = std::abs(_max_velocities[0] / _accel_limits[0]);
    for (int i = 0; i < 3; ++i) max_accel_time = std::max(max_accel_time, std::abs(_max_velocities[i] / _accel_limits[i]));
    if (max_accel_time != 0.0f) {
        float inv = 1.0f / max_accel_time;
        for (int i = 0; i < 3; ++i) _accelerations[i] = inv * _max_velocities[i];
    } else {
        for (int i = 0; i < 3; ++i) _accelerations[i] = 0.0f;
    }
    _acceleration_duration = max_accel_time;
}

void RampVelocityControlSpace::nullControl(float p[5]) const
{
    for (int i = 0; i < 5; ++i) p[i] = 0.0f;
}

void RampVelocityControlSpace::sample(util::random::RNG& rng, float p[5]) const
{
    bool sampled[3] = {false, false, false};
    if (_xy_ball) {
        double buf[2];
        rng.uniformInBall(_xy_norm_limit, 2, buf);
        p[0] = (float)buf[0];
        p[1] = (float)buf[1];
        sampled[0] = sampled[1] = true;
    }
    for (int i = 0; i < 3; ++i)
        if (!sampled[i]) p[i] = (float)rng.uniformReal(-_limits.velocity_limits[i], _limits.velocity_limits[i]);
    p[3] = (float)rng.uniformReal(_limits.duration_limits[0], _limits.duration_limits[1]);
    p[4] = 0.0f;
}

bool SimEnvStatePropagator::propagate(const state::SimEnvWorldState* state, float control[5],
                                      state::SimEnvWorldState* result) const
{
    int32_t ok = 0;
    std::vector<float> in(state->values);  // in / out may alias (OraclePushPlanner.cpp:510-512)
    result->values.resize(in.size());
    _ctx->check(mps_propagate(_ctx->get(), in.data(), control, result->values.data(), &ok, nullptr),
                "mps::planner::ompl::control::SimEnvStatePropagator::propagate");
    return ok != 0;
}

NaiveControlSampler::NaiveControlSampler(DeviceContextPtr ctx, RampVelocityControlSpacePtr control_space,
                                         state::SimEnvWorldStateDistanceMeasurePtr distance, util::random::RNGPtr rng,
                                         unsigned int k)
    : _ctx(ctx), _control_space(control_space), _distance(distance), _rng(rng), _k(k > 0 ? k : 1)
{
}

unsigned int NaiveControlSampler::sampleTo(float control[5], const state::SimEnvWorldState* source,
                                           state::SimEnvWorldState* dest)
{
    _controls.resize((size_t)_k * 5);
    for (unsigned int i = 0; i < _k; ++i) _control_space->sample(*_rng, _controls.data() + 5 * i);
    mps_extend_in in;
    std::memset(&in, 0, sizeof(in));
    in.start = source->values.data();
    in.controls = _controls.data();
    in.K = (int32_t)_k;
    in.L = 1;
    in.dest = dest->values.data();
    in.weights = _distance->weights().data();
    in.mask = _distance->mask();
    mps_extend_best best;
    std::vector<float> best_state(source->values.size());
    float best_control[5];
    uint32_t best_flags = 0;
    mps_extend_out out;
    std::memset(&out, 0, sizeof(out));
    out.best = &best;
    out.best_states = best_state.data();
    out.best_controls = best_control;
    out.best_flags = &best_flags;
    _ctx->check(mps_extend(_ctx->get(), &in, &out), "mps::planner::ompl::control::NaiveControlSampler::sampleTo");
    if (best.k < 0) {
        _control_space->nullControl(control);
        dest->values = source->values;
        return 0;
    }
    for (int i = 0; i < 5; ++i) control[i] = best_control[i];
    dest->values = best_state;
    return 1;
}
}  // namespace control

namespace planning {
namespace essentials {
void Path::initBacktrackMotion(MotionPtr last)
{
    motions.clear();
    for (MotionPtr m = last; m; m = m->parent) motions.push_back(m);
    std::reverse(motions.begin(), motions.end());
}
}  // namespace essentials
}  // namespace planning
}  // namespace ompl

// ========================================================================= DeviceNearestNeighbors
using ompl::planning::essentials::MotionPtr;

void DeviceNearestNeighbors::clear()
{
    _ctx->check(mps_tree_clear(_ctx->get()), "DeviceNearestNeighbors::clear");
    _motions.clear();
}

void DeviceNearestNeighbors::add(const MotionPtr& m, int32_t slice_id)
{
    int64_t idx = -1;
    int32_t parent = m->parent ? (int32_t)m->parent->index : -1;
    _ctx->check(mps_tree_add(_ctx->get(), m->state.values.data(), parent, m->control.data(), slice_id, &idx),
                "DeviceNearestNeighbors::add");
    if (idx != (int64_t)_motions.size()) throw std::logic_error("[DeviceNearestNeighbors::add] host / device tree out of step");
    m->index = idx;
    _motions.push_back(m);
}

void DeviceNearestNeighbors::adopt(const MotionPtr& m, int64_t index)
{
    if (index != (int64_t)_motions.size()) throw std::logic_error("[DeviceNearestNeighbors::adopt] host / device tree out of step");
    m->index = index;
    _motions.push_back(m);
}

MotionPtr DeviceNearestNeighbors::nearest(const ompl::state::SimEnvWorldState& query, const std::vector<float>& weights,
                                          uint32_t mask, int32_t slice_filter, double* distance) const
{
    int64_t idx = -1;
    double d = 0.0;
    _ctx->check(mps_tree_nearest(_ctx->get(), query.values.data(), weights.empty() ? nullptr : weights.data(), mask,
                                 slice_filter, &idx, &d),
                "DeviceNearestNeighbors::nearest");
    if (distance) *distance = d;
    if (idx < 0) return nullptr;
    return _motions.at((size_t)idx);
}

// ========================================================================================= oracle
namespace pushing {
namespace oracle {
void RampComputer::steer(const float cur[3], const float des[3], std::vector<std::array<float, 5>>& control_params) const
{
    // computeDirection of the robot configuration space: R^2 difference + SO(2) shortest direction
    float dir[3] = {des[0] - cur[0], des[1] - cur[1], util::math::shortest_direction_so2(cur[2], des[2])};
    float distance = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    if (!(distance > 0.0f)) return;
    for (int i = 0; i < 3; ++i) dir[i] /= distance;
    const auto& lim = _control_space->limits();
    float vabs = std::numeric_limits<float>::max();
    for (int i = 0; i < 3; ++i) vabs = std::min(vabs, lim.velocity_limits[i] / std::abs(dir[i]));
    float vel[3] = {vabs * dir[0], vabs * dir[1], vabs * dir[2]};
    ompl::control::RampVelocityControl ramp(lim.acceleration_limits);
    ramp.setMaxVelocities(vel, 0.0f);
    float accel_time = ramp.getAccelerationTime();
    auto norm3 = [](const float v[3], float s) { return std::sqrt(s * v[0] * s * v[0] + s * v[1] * s * v[1] + s * v[2] * s * v[2]); };
    float dist_accel = norm3(vel, accel_time);
    int guard = 0;
    while (distance > 0.0f && guard++ < 64) {
        if (dist_accel > distance) {
            vabs = distance / accel_time;
            for (int i = 0; i < 3; ++i) vel[i] = vabs * dir[i];
            ramp.setMaxVelocities(vel, 0.0f);
            accel_time = ramp.getAccelerationTime();
            dist_accel = norm3(vel, accel_time);
        }
        distance = distance - dist_accel;
        float max_plateau_dist = std::min(norm3(vel, lim.duration_limits[1]), distance);
        float duration = max_plateau_dist / vabs;
        ramp.setMaxVelocities(vel, duration);
        std::array<float, 5> p;
        ramp.getParameters(p.data());
        control_params.push_back(p);
        distance = distance - max_plateau_dist;
    }
}

HumanOracle::HumanOracle(RobotOraclePtr robot_oracle, const std::vector<ObjectData>& object_data,
                         util::random::RNGPtr rng, const Parameters& params)
    : _params(params), _robot_steerer(robot_oracle), _object_data(object_data), _rng(rng)
{
}

void HumanOracle::relativeSE2(const float state[3], const float ref[3], float out[3])
{
    out[0] = state[0] - ref[0];
    out[1] = state[1] - ref[1];
    out[2] = util::math::shortest_direction_so2(ref[2], state[2]);
}

float HumanOracle::predictPushability(const float cur_obj[3], const float next_obj[3]) const
{
    float rel[3];
    relativeSE2(next_obj, cur_obj, rel);
    float maha = 0.0f;
    for (int i = 0; i < 3; ++i) maha += rel[i] * rel[i] / _params.pushability_covariance[i];
    maha = std::sqrt(maha);
    if (maha == 0.0f) return std::numeric_limits<float>::max();
    return 1.0f / maha;
}

float HumanOracle::predictFeasibility(const float cur_robot[3], const float cur_obj[3], const float next_obj[3]) const
{
    float obj_diff[3], rel_robot[3];
    relativeSE2(next_obj, cur_obj, obj_diff);
    relativeSE2(cur_robot, cur_obj, rel_robot);
    float push[2] = {obj_diff[0], obj_diff[1]};
    float appr[2] = {-rel_robot[0], -rel_robot[1]};
    float robot_dist = std::sqrt(appr[0] * appr[0] + appr[1] * appr[1]);
    float change = std::sqrt(push[0] * push[0] + push[1] * push[1]);
    if (robot_dist == 0.0f) return 0.0f;
    float approach_angle = 0.0f;
    if (change != 0.0f) approach_angle = std::acos((appr[0] * push[0] + appr[1] * push[1]) / (robot_dist * change));
    float args = std::pow(robot_dist - _params.optimal_push_distance, 2.0f) / (2.0f * std::pow(_params.push_distance_tolerance, 2.0f));
    float distance_feasibility = std::exp(-args);
    args = std::pow(approach_angle, 2.0f) / (2.0f * std::pow(_params.push_angle_tolerance, 2.0f));
    return distance_feasibility * std::exp(-args);
}

void HumanOracle::predictAction(const float cur_robot[3], const float cur_obj[3], const float next_obj[3],
                                unsigned int obj_id, float control[5]) const
{
    float rel[3];
    relativeSE2(next_obj, cur_obj, rel);
    float next_robot[3] = {next_obj[0], next_obj[1], cur_robot[2]};
    float n = std::sqrt(rel[0] * rel[0] + rel[1] * rel[1]);
    for (int i = 0; i < 5; ++i) control[i] = 0.0f;
    if (n != 0.0f) {
        float heading[2] = {rel[0] / n, rel[1] / n};
        float obj_size = std::fmax(_object_data.at(obj_id).width, _object_data.at(obj_id).height);
        next_robot[0] -= heading[0] * obj_size / 2.0f;
        next_robot[1] -= heading[1] * obj_size / 2.0f;
        std::vector<std::array<float, 5>> controls;
        _robot_steerer->steer(cur_robot, next_robot, controls);
        if (!controls.empty())
            for (int i = 0; i < 5; ++i) control[i] = controls[0][i];
    }
}

void HumanOracle::sampleFeasibleState(const float cur_obj[3], const float next_obj[3], unsigned int, float new_robot[3])
{
    double pushing_dist = _rng->gaussian(_params.optimal_push_distance, _params.push_distance_tolerance);
    double angle = _rng->gaussian(0.0, _params.push_angle_tolerance);
    float rel[3];
    relativeSE2(next_obj, cur_obj, rel);
    if (std::sqrt(rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]) == 0.0f) {
        rel[0] = 0.0001f * (0.5f - (float)_rng->uniform01());
        rel[1] = 0.0001f * (0.5f - (float)_rng->uniform01());
    }
    float n = std::sqrt(rel[0] * rel[0] + rel[1] * rel[1]);
    float pd[2] = {1.0f, 0.0f};
    if (n > 0.0f) {
        pd[0] = rel[0] / n;
        pd[1] = rel[1] / n;
    }
    float ca = std::cos((float)angle), sa = std::sin((float)angle);
    float off[2] = {(float)pushing_dist * (ca * pd[0] - sa * pd[1]), (float)pushing_dist * (sa * pd[0] + ca * pd[1])};
    new_robot[0] = cur_obj[0] - off[0];
    new_robot[1] = cur_obj[1] - off[1];
    new_robot[2] = std::atan2(pd[1], pd[0]) - 1.57f;  // specific to the reference's floating end-effector
}

float HumanOracle::getMaximalPushingDistance() const
{
    return std::sqrt(2.0f * _params.pushability_covariance[0] + 2.0f * _params.pushability_covariance[1]) +
           0.01f * std::sqrt(_params.pushability_covariance[2]);
}

OracleControlSampler::OracleControlSampler(ompl::state::SimEnvWorldStateSpacePtr space,
                                           ompl::control::RampVelocityControlSpacePtr cs, PushingOraclePtr pushing_oracle,
                                           RobotOraclePtr robot_oracle, unsigned int robot_id, util::random::RNGPtr rng)
    : _space(space), _control_space(cs), _pushing_oracle(pushing_oracle), _robot_oracle(robot_oracle), _robot_id(robot_id),
      _rng(rng)
{
}

void OracleControlSampler::sampleFeasibleState(ompl::state::SimEnvWorldState* x_state,
                                               const ompl::state::SimEnvWorldState* x_prime_state,
                                               unsigned int local_target_obj, float p_uniform)
{
    float* x_r = x_state->getObjectState(_robot_id);
    if (p_uniform > 0.0f) {
        float die = (float)_rng->uniform01();
        if (die <= p_uniform) {
            _space->sampleUniformObject(*_rng, x_r);
            return;
        }
    }
    float new_robot[3] = {x_r[0], x_r[1], x_r[2]};
    _pushing_oracle->sampleFeasibleState(x_state->getObjectState(local_target_obj),
                                         x_prime_state->getObjectState(local_target_obj), local_target_obj, new_robot);
    x_r[0] = new_robot[0];
    x_r[1] = new_robot[1];
    double th = (double)new_robot[2];  // setConfiguration normalises the SO(2) component
    util::math::normalize_orientation(th);
    x_r[2] = (float)th;
}

bool OracleControlSampler::steerRobot(std::vector<std::array<float, 5>>& controls,
                                      const ompl::state::SimEnvWorldState* source,
                                      const ompl::state::SimEnvWorldState* dest)
{
    size_t before = controls.size();
    _robot_oracle->steer(source->getObjectState(_robot_id), dest->getObjectState(_robot_id), controls);
    return controls.size() > before;
}

bool OracleControlSampler::steerPush(std::vector<std::array<float, 5>>& controls,
                                     const ompl::state::SimEnvWorldState* source,
                                     const ompl::state::SimEnvWorldState* dest, unsigned int obj_id, float action_noise)
{
    if (action_noise > 0.0f) {
        float die = (float)_rng->uniform01();
        if (die <= action_noise) {
            randomControl(controls);
            return true;
        }
    }
    std::array<float, 5> c;
    _pushing_oracle->predictAction(source->getObjectState(_robot_id), source->getObjectState(obj_id),
                                   dest->getObjectState(obj_id), obj_id, c.data());
    controls.push_back(c);
    return true;
}

void OracleControlSampler::randomControl(std::vector<std::array<float, 5>>& controls)
{
    std::array<float, 5> c;
    _control_space->sample(*_rng, c.data());
    controls.push_back(c);
}
}  // namespace oracle

// ====================================================================================== algorithm
namespace algorithm {
void ParamSet::declareParam(const std::string& name, std::function<void(const std::string&)> setter,
                            std::function<std::string()> getter)
{
    _setters[name] = setter;
    _getters[name] = getter;
}

bool ParamSet::setParam(const std::string& name, const std::string& value)
{
    auto it = _setters.find(name);
    if (it == _setters.end()) return false;
    it->second(value);
    return true;
}

bool ParamSet::getParam(const std::string& name, std::string& value) const
{
    auto it = _getters.find(name);
    if (it == _getters.end()) return false;
    value = it->second();
    return true;
}

RearrangementRRT::RearrangementRRT(SpaceInformationPtr si) : _si(si)
{
    _state_space = si->state_space;
    if (!_state_space) throw std::logic_error("[mps::planner::pushing::algorithm::RearrangementRRT::setup] Could not cast state space to SimEnvWorldStateSpace");
    _distance_measure = std::make_shared<ompl::state::SimEnvWorldStateDistanceMeasure>(_state_space);
    _tree = std::make_shared<DeviceNearestNeighbors>(si->ctx);
    _rng = si->rng;
    if (_state_space->getNumObjects() <= 1) throw std::logic_error("RearrangementRRT needs a robot and at least one object");
}

uint32_t RearrangementRRT::fullMask() const
{
    unsigned int B = _state_space->getNumObjects();
    return B < 32 ? (1u << B) - 1u : 0xffffffffu;
}

void RearrangementRRT::setup(PlanningQueryPtr pq, PlanningBlackboard& blackboard)
{
    blackboard.robot_id = pq->robot_id;
    _distance_measure->setAll(true);
    _distance_measure->setWeights(pq->weights);
    _tree->clear();
}

bool RearrangementRRT::plan(PlanningQueryPtr pq)
{
    PlanningStatistics stats;
    return plan(pq, stats);
}

bool RearrangementRRT::plan(PlanningQueryPtr pq, PlanningStatistics& stats)
{
    PlanningBlackboard blackboard(pq);
    setup(pq, blackboard);
    auto wall0 = std::chrono::steady_clock::now();
    MotionPtr current_motion = getNewMotion();
    MotionPtr sample_motion = getNewMotion();
    MotionPtr final_motion = nullptr;
    current_motion->state = pq->start_state;
    _si->control_space->nullControl(current_motion->control.data());
    addToTree(current_motion, nullptr, blackboard);
    final_motion = current_motion;
    bool solved = pq->goal_region->isSatisfied(&current_motion->state);
    timer_ptr->startTimer(pq->time_out);
    while (!timer_ptr->timeOutExceeded() && !pq->stopping_condition() && !solved) {
        if (pq->max_iterations && blackboard.stats.num_iterations >= pq->max_iterations) break;
        blackboard.stats.num_iterations++;
        unsigned int active_obj_id = 0;
        bool goal_sampled = sample(sample_motion, active_obj_id, blackboard);
        selectTreeNode(sample_motion, current_motion, active_obj_id, goal_sampled, blackboard);
        solved = extend(current_motion, &sample_motion->state, active_obj_id, final_motion, blackboard);
    }
    blackboard.stats.runtime = timer_ptr->stopTimer();
    blackboard.stats.wall_time = std::chrono::duration<float>(std::chrono::steady_clock::now() - wall0).count();
    blackboard.stats.success = solved;
    if (solved) {
        pq->path = std::make_shared<ompl::planning::essentials::Path>();
        pq->path->initBacktrackMotion(final_motion);
    }
    stats = blackboard.stats;
    return solved;
}

PlanningQueryPtr RearrangementRRT::createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                                       const ompl::state::SimEnvWorldState& start_state,
                                                       unsigned int robot_id, float timeout)
{
    auto pq = std::make_shared<PlanningQuery>();
    pq->goal_region = goal_region;
    pq->start_state = start_state;
    pq->robot_id = robot_id;
    pq->time_out = timeout;
    pq->parameters->declareParam("goal_bias", [this](const std::string& s) { setGoalBias(std::stof(s)); }, [this]() { return f2s(getGoalBias()); });
    pq->parameters->setParam("goal_bias", "0.2");
    pq->parameters->declareParam("target_bias", [this](const std::string& s) { setTargetBias(std::stof(s)); }, [this]() { return f2s(getTargetBias()); });
    pq->parameters->setParam("target_bias", "0.25");
    pq->parameters->declareParam("robot_bias", [this](const std::string& s) { setRobotBias(std::stof(s)); }, [this]() { return f2s(getRobotBias()); });
    pq->parameters->setParam("robot_bias", "0.0");
    return pq;
}

void RearrangementRRT::sampleValidState(ompl::state::SimEnvWorldState* s)
{
    // ompl::base::UniformValidStateSampler: uniform samples until one is valid (default 100 attempts);
    // candidates are validated 8 at a time by one device launch
    const unsigned int per_round = 8, max_attempts = 100;
    std::vector<ompl::state::SimEnvWorldState> cands;
    std::vector<uint8_t> valid;
    for (unsigned int attempts = 0; attempts < max_attempts; attempts += per_round) {
        cands.assign(per_round, ompl::state::SimEnvWorldState(_state_space->getNumObjects()));
        for (auto& c : cands) _state_space->sampleUniform(*_rng, &c);
        _si->validity_checker->isValidBatch(cands, valid);
        for (size_t i = 0; i < cands.size(); ++i) {
            if (valid[i]) {
                *s = cands[i];
                return;
            }
        }
    }
    *s = cands.back();
}

bool RearrangementRRT::sample(MotionPtr motion, unsigned int& target_obj_id, PlanningBlackboard& pb)
{
    bool is_goal = false;
    if (_rng->uniform01() < _goal_bias && pb.pq->goal_region->canSample()) {
        pb.pq->goal_region->sampleGoal(&motion->state);
        target_obj_id = pb.pq->goal_region->sampleTargetObjectIndex();
        is_goal = true;
    } else {
        sampleValidState(&motion->state);
        target_obj_id = sampleActiveObject(pb);
    }
    pb.stats.num_samples++;
    return is_goal;
}

void RearrangementRRT::selectTreeNode(const MotionPtr& sample_motion, MotionPtr& selected_node, unsigned int&, bool,
                                      PlanningBlackboard& pb)
{
    _distance_measure->setAll(true);  // the full state counts here (RRT.cpp:201)
    selected_node = _tree->nearest(sample_motion->state, _distance_measure->weights(), _distance_measure->mask());
    pb.stats.num_nearest_neighbor_queries++;
}

void RearrangementRRT::addToTree(MotionPtr new_motion, MotionPtr parent, PlanningBlackboard&)
{
    new_motion->parent = parent;
    _tree->add(new_motion);
}

unsigned int RearrangementRRT::sampleActiveObject(const PlanningBlackboard& pb) const
{
    double random_value = _rng->uniform01();
    if (random_value < _target_bias) {
        return pb.pq->goal_region->sampleTargetObjectIndex();
    } else if (random_value < _target_bias + _robot_bias) {
        return pb.robot_id;
    }
    return (unsigned int)_rng->uniformInt(0, (int)_state_space->getNumObjects() - 1);
}

// ---- NaiveRearrangementRRT ---------------------------------------------------------------------
NaiveRearrangementRRT::NaiveRearrangementRRT(SpaceInformationPtr si, unsigned int k) : RearrangementRRT(si), _num_control_samples(k) {}

void NaiveRearrangementRRT::setup(PlanningQueryPtr pq, PlanningBlackboard& blackboard) { RearrangementRRT::setup(pq, blackboard); }

PlanningQueryPtr NaiveRearrangementRRT::createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                                            const ompl::state::SimEnvWorldState& start_state,
                                                            unsigned int robot_id, float timeout)
{
    auto pq = RearrangementRRT::createPlanningQuery(goal_region, start_state, robot_id, timeout);
    pq->parameters->declareParam("num_control_samples", [this](const std::string& s) { setNumControlSamples((unsigned int)std::stoul(s)); },
                                 [this]() { return std::to_string(getNumControlSamples()); });
    pq->parameters->setParam("num_control_samples", "10");
    return pq;
}

bool NaiveRearrangementRRT::extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int active_obj_id,
                                   MotionPtr& last_motion, PlanningBlackboard& pb)
{
    // The K-sample loop of NaiveControlSampler::sampleTo, the goal test (:377) and addToTree (:383) are one
    // device extend: K controls in, the winner appended to the device tree, one record out.
    _distance_measure->setAll(false);
    _distance_measure->setActive(active_obj_id, true);
    last_motion = start;
    const unsigned int K = _num_control_samples > 0 ? _num_control_samples : 1;
    _controls.resize((size_t)K * 5);
    for (unsigned int i = 0; i < K; ++i) _si->control_space->sample(*_rng, _controls.data() + 5 * i);
    mps_extend_in in;
    std::memset(&in, 0, sizeof(in));
    in.start = start->state.values.data();
    in.controls = _controls.data();
    in.K = (int32_t)K;
    in.L = 1;
    in.dest = dest->values.data();
    in.weights = _distance_measure->weights().data();
    in.mask = _distance_measure->mask();
    in.goals = pb.pq->goal_region->abiGoals().data();
    in.n_goals = (int32_t)pb.pq->goal_region->abiGoals().size();
    in.append_to_tree = 1;
    in.parent = (int32_t)start->index;
    MotionPtr new_motion = getNewMotion();
    mps_extend_best best;
    uint32_t flags = 0;
    mps_extend_out out;
    std::memset(&out, 0, sizeof(out));
    out.best = &best;
    out.best_states = new_motion->state.values.data();
    out.best_controls = new_motion->control.data();
    out.best_flags = &flags;
    _si->ctx->check(mps_extend(_si->ctx->get(), &in, &out), "NaiveRearrangementRRT::extend");
    pb.stats.num_state_propagations += K;
    bool reached_a_goal = false;
    if (best.k >= 0) {
        new_motion->parent = start;
        _tree->adopt(new_motion, best.tree_index);
        last_motion = new_motion;
        reached_a_goal = (flags & MPS_F_GOAL) != 0;
    } else {
        // no valid control: sampleTo leaves dest := source; the goal test then runs on the start state
        reached_a_goal = pb.pq->goal_region->isSatisfied(&start->state);
    }
    return reached_a_goal;
}

// ---- HybridActionRRT ---------------------------------------------------------------------------
HybridActionRRT::HybridActionRRT(SpaceInformationPtr si, oracle::PushingOraclePtr pushing_oracle,
                                 oracle::RobotOraclePtr robot_oracle, unsigned int robot_id)
    : RearrangementRRT(si)
{
    _oracle_sampler = std::make_shared<oracle::OracleControlSampler>(si->state_space, si->control_space, pushing_oracle,
                                                                     robot_oracle, robot_id, si->rng);
}

PlanningQueryPtr HybridActionRRT::createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                                      const ompl::state::SimEnvWorldState& start_state,
                                                      unsigned int robot_id, float timeout)
{
    auto pq = RearrangementRRT::createPlanningQuery(goal_region, start_state, robot_id, timeout);
    pq->parameters->declareParam("num_control_samples", [this](const std::string& s) { setNumControlSamples((unsigned int)std::stoul(s)); },
                                 [this]() { return std::to_string(getNumControlSamples()); });
    pq->parameters->setParam("num_control_samples", "10");
    pq->parameters->declareParam("action_randomness", [this](const std::string& s) { setActionRandomness(std::stof(s)); },
                                 [this]() { return f2s(getActionRandomness()); });
    pq->parameters->setParam("action_randomness", "0.0001");
    return pq;
}

void HybridActionRRT::sampleActionSequence(std::vector<std::array<float, 5>>& controls, MotionPtr start,
                                           ompl::state::SimEnvWorldState* dest, PlanningBlackboard& pb)
{
    float random_die = (float)_rng->uniform01();
    if (random_die < _action_randomness) {
        _oracle_sampler->randomControl(controls);
    } else {
        unsigned int obj_id = sampleActiveObject(pb);
        if (obj_id == pb.robot_id) {  // transit primitive
            _oracle_sampler->steerRobot(controls, &start->state, dest);
        } else {  // transfer primitive
            ompl::state::SimEnvWorldState feasible(start->state);
            _oracle_sampler->sampleFeasibleState(&feasible, dest, obj_id);
            _oracle_sampler->steerRobot(controls, &start->state, &feasible);
            // the robot is "beamed" to the feasible state for the push prediction (RRT.cpp:521-525)
            _oracle_sampler->steerPush(controls, &feasible, dest, obj_id);
        }
    }
}

bool HybridActionRRT::extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int, MotionPtr& last_motion,
                             PlanningBlackboard& pb)
{
    // K action sequences are generated on the host (oracles), propagated as K chains by one launch;
    // argmin on (float) distance with the full mask (set by selectTreeNode), best chain appended on device.
    const unsigned int K = _num_control_samples > 0 ? _num_control_samples : 1;
    std::vector<std::vector<std::array<float, 5>>> seqs(K);
    int L = 1;
    for (unsigned int i = 0; i < K; ++i) {
        sampleActionSequence(seqs[i], start, dest, pb);
        if ((int)seqs[i].size() > kMaxChain) seqs[i].resize(kMaxChain);
        L = std::max(L, (int)seqs[i].size());
    }
    std::vector<float> controls((size_t)K * L * 5, 0.0f);
    std::vector<int32_t> n_ctrl(K);
    for (unsigned int i = 0; i < K; ++i) {
        n_ctrl[i] = (int32_t)seqs[i].size();
        for (size_t l = 0; l < seqs[i].size(); ++l)
            for (int c = 0; c < 5; ++c) controls[((size_t)i * L + l) * 5 + c] = seqs[i][l][c];
    }
    mps_extend_in in;
    std::memset(&in, 0, sizeof(in));
    in.start = start->state.values.data();
    in.controls = controls.data();
    in.n_ctrl = n_ctrl.data();
    in.K = (int32_t)K;
    in.L = L;
    in.dest = dest->values.data();
    in.weights = _distance_measure->weights().data();
    in.mask = _distance_measure->mask();
    in.compare_f32 = 1;
    in.goals = pb.pq->goal_region->abiGoals().data();
    in.n_goals = (int32_t)pb.pq->goal_region->abiGoals().size();
    in.append_to_tree = 1;
    in.parent = (int32_t)start->index;
    const size_t dim = start->state.values.size();
    std::vector<float> best_states((size_t)L * dim), best_controls((size_t)L * 5);
    std::vector<uint32_t> best_flags(L);
    std::vector<int32_t> all_steps((size_t)K * L);
    std::vector<uint32_t> all_flags((size_t)K * L);
    mps_extend_best best;
    mps_extend_out out;
    std::memset(&out, 0, sizeof(out));
    out.best = &best;
    out.best_states = best_states.data();
    out.best_controls = best_controls.data();
    out.best_flags = best_flags.data();
    out.all_flags = all_flags.data();
    _si->ctx->check(mps_extend(_si->ctx->get(), &in, &out), "HybridActionRRT::extend");
    for (uint32_t f : all_flags)
        if (f & MPS_F_RAN) pb.stats.num_state_propagations++;
    if (best.k < 0) return false;
    int n_add = best.goal_step >= 0 ? best.goal_step + 1 : best.n_ok;
    MotionPtr prev = start;
    for (int l = 0; l < n_add; ++l) {
        MotionPtr m = getNewMotion();
        std::memcpy(m->state.values.data(), best_states.data() + (size_t)l * dim, dim * sizeof(float));
        for (int c = 0; c < 5; ++c) m->control[c] = best_controls[(size_t)l * 5 + c];
        m->parent = prev;
        _tree->adopt(m, best.tree_index + l);
        prev = m;
        last_motion = m;
    }
    return best.goal_step >= 0;
}

// ---- OracleRearrangementRRT --------------------------------------------------------------------
OracleRearrangementRRT::OracleRearrangementRRT(SpaceInformationPtr si, oracle::PushingOraclePtr pushing_oracle,
                                               oracle::RobotOraclePtr robot_oracle, unsigned int robot_id)
    : RearrangementRRT(si)
{
    _oracle_sampler = std::make_shared<oracle::OracleControlSampler>(si->state_space, si->control_space, pushing_oracle,
                                                                     robot_oracle, robot_id, si->rng);
}

PlanningQueryPtr OracleRearrangementRRT::createPlanningQuery(ompl::state::goal::ObjectsRelocationGoalPtr goal_region,
                                                             const ompl::state::SimEnvWorldState& start_state,
                                                             unsigned int robot_id, float timeout)
{
    auto pq = RearrangementRRT::createPlanningQuery(goal_region, start_state, robot_id, timeout);
    pq->parameters->declareParam("action_randomness", [this](const std::string& s) { setActionRandomness(std::stof(s)); },
                                 [this]() { return f2s(getActionRandomness()); });
    pq->parameters->setParam("action_randomness", "0.0001");
    pq->parameters->declareParam("state_noise", [this](const std::string& s) { setStateNoise(std::stof(s)); },
                                 [this]() { return f2s(getStateNoise()); });
    pq->parameters->setParam("state_noise", "0.00001");
    return pq;
}

void OracleRearrangementRRT::selectTreeNode(const MotionPtr& sample_motion, MotionPtr& selected_node,
                                            unsigned int& active_obj_id, bool sample_is_goal, PlanningBlackboard& pb)
{
    RearrangementRRT::selectTreeNode(sample_motion, selected_node, active_obj_id, sample_is_goal, pb);
    if (active_obj_id != pb.robot_id) {
        ompl::state::SimEnvWorldState tmp(selected_node->state);
        _oracle_sampler->sampleFeasibleState(&tmp, &sample_motion->state, active_obj_id, _feasible_state_noise);
        _state_space->copySubState(&sample_motion->state, &tmp, pb.robot_id);
    }
}

bool OracleRearrangementRRT::extend(MotionPtr start, ompl::state::SimEnvWorldState* dest, unsigned int active_obj_id,
                                    MotionPtr& last_motion, PlanningBlackboard& pb)
{
    std::vector<std::array<float, 5>> controls;
    bool b_goal = false, extension_success = false;
    float random_die = (float)_rng->uniform01();
    if (random_die < _action_randomness) {
        _oracle_sampler->randomControl(controls);
        extendStep(controls, start, last_motion, pb, extension_success, b_goal);
    } else {
        _oracle_sampler->steerRobot(controls, &start->state, dest);
        if (controls.empty()) return false;  // "OracleControlSampler provided no controls at all"
        extendStep(controls, start, last_motion, pb, extension_success, b_goal);
        if (extension_success && active_obj_id != pb.robot_id && !b_goal) {
            controls.clear();
            _oracle_sampler->steerPush(controls, &last_motion->state, dest, active_obj_id);
            extendStep(controls, last_motion, last_motion, pb, extension_success, b_goal);
        }
    }
    return b_goal;
}

void OracleRearrangementRRT::extendStep(const std::vector<std::array<float, 5>>& controls_in, const MotionPtr& start_motion,
                                        MotionPtr& result_motion, PlanningBlackboard& pb, bool& extension_success,
                                        bool& goal_reached)
{
    // One chain of controls = one device launch (K = 1); every successful step becomes a tree node, the
    // chain stops at the first failure or the first goal state (RRT.cpp:698-735).
    goal_reached = false;
    extension_success = false;
    std::vector<std::array<float, 5>> controls(controls_in);
    if (controls.empty()) return;
    if (controls.size() > 16) controls.resize(16);
    const int L = (int)controls.size();
    std::vector<float> flat((size_t)L * 5);
    for (int l = 0; l < L; ++l)
        for (int c = 0; c < 5; ++c) flat[(size_t)l * 5 + c] = controls[l][c];
    mps_extend_in in;
    std::memset(&in, 0, sizeof(in));
    in.start = start_motion->state.values.data();
    in.controls = flat.data();
    in.K = 1;
    in.L = L;
    in.goals = pb.pq->goal_region->abiGoals().data();
    in.n_goals = (int32_t)pb.pq->goal_region->abiGoals().size();
    const size_t dim = start_motion->state.values.size();
    std::vector<float> states((size_t)L * dim), rest(L);
    std::vector<uint32_t> flags(L);
    mps_extend_out out;
    std::memset(&out, 0, sizeof(out));
    out.all_states = states.data();
    out.all_rest = rest.data();
    out.all_flags = flags.data();
    _si->ctx->check(mps_extend(_si->ctx->get(), &in, &out), "OracleRearrangementRRT::extendStep");
    MotionPtr prev = start_motion;
    for (int l = 0; l < L; ++l) {
        if (!(flags[l] & MPS_F_RAN)) break;
        pb.stats.num_state_propagations++;
        extension_success = (flags[l] & MPS_F_OK) != 0;
        if (!extension_success) return;
        MotionPtr m = getNewMotion();
        std::memcpy(m->state.values.data(), states.data() + (size_t)l * dim, dim * sizeof(float));
        for (int c = 0; c < 4; ++c) m->control[c] = controls[l][c];
        m->control[4] = rest[l];
        addToTree(m, prev, pb);
        result_motion = m;
        if (flags[l] & MPS_F_GOAL) {
            goal_reached = true;
            return;
        }
        prev = m;
    }
}

// ---- SliceBasedOracleRRT -----------------------------------------------------------------------
SliceBasedOracleRRT::SliceBasedOracleRRT(SpaceInformationPtr si, oracle::PushingOraclePtr pushing_oracle,
                                         oracle::RobotOraclePtr robot_oracle, unsigned int robot_id)
    : OracleRearrangementRRT(si, pushing_oracle, robot_oracle, robot_id), _pushing_oracle(pushing_oracle)
{
    // the slice representatives live in their own device tree (index = slice id)
    auto slice_ctx = std::make_shared<DeviceContext>(si->ctx->scene(), si->ctx->device());
    _slices_nn = std::make_shared<DeviceNearestNeighbors>(slice_ctx);
}

uint32_t SliceBasedOracleRRT::arrangementMask(unsigned int robot_id) const { return fullMask() & ~(1u << robot_id); }

void SliceBasedOracleRRT::setup(PlanningQueryPtr pq, PlanningBlackboard& pb)
{
    RearrangementRRT::setup(pq, pb);
    _slices_nn->clear();
    _slices_list.clear();
    _robot_id_cached = pb.robot_id;
    _min_slice_distance = (_state_space->getNumObjects() - 1) * 0.001;  // RRT.cpp:780
}

SlicePtr SliceBasedOracleRRT::getSlice(const ompl::state::SimEnvWorldState& state, PlanningBlackboard& pb,
                                       double* distance) const
{
    if (_slices_nn->size() == 0) return nullptr;
    // ObjectArrangementDistanceFn: every object but the robot (algorithm/Interfaces.cpp:88-100)
    MotionPtr rep = _slices_nn->nearest(state, _distance_measure->weights(), arrangementMask(pb.robot_id), -1, distance);
    ++pb.stats.num_nearest_neighbor_queries;
    return rep ? _slices_list.at((size_t)rep->index) : nullptr;
}

bool SliceBasedOracleRRT::projectSliceBall(const ompl::state::SimEnvWorldStateDistanceMeasure& measure, unsigned int n_objects,
                                           unsigned int robot_id, float radius,
                                           const ompl::state::SimEnvWorldState& center,
                                           ompl::state::SimEnvWorldState* sample)
{
    // d_arr(sample, center) <= r: inside the ball; else move every object towards the centre's arrangement by the
    // factor r / d along the straight line (positions) and the shortest arc (angles); the robot's sub-state stays
    auto dm = measure;
    dm.setAll(true);
    dm.setActive(robot_id, false);
    double d = dm.distance(sample, &center);
    if (d <= (double)radius || d == 0.0) return true;
    float f = (float)((double)radius / d);
    for (unsigned int i = 0; i < n_objects; ++i) {
        if (i == robot_id) continue;
        const float* c = center.getObjectState(i);
        float* s = sample->getObjectState(i);
        s[0] = c[0] + f * (s[0] - c[0]);
        s[1] = c[1] + f * (s[1] - c[1]);
        float th = c[2] + f * util::math::shortest_direction_so2(c[2], s[2]);
        util::math::normalize_orientation(th);
        s[2] = th;
    }
    return false;
}

bool SliceBasedOracleRRT::projectSliceBall(const ompl::state::SimEnvWorldState& center,
                                           ompl::state::SimEnvWorldState* sample) const
{
    return projectSliceBall(*_distance_measure, _state_space->getNumObjects(), _robot_id_cached, _slice_ball_radius, center,
                            sample);
}

void SliceBasedOracleRRT::selectTreeNode(const MotionPtr& sample, MotionPtr& selected_node, unsigned int& active_obj_id,
                                         bool, PlanningBlackboard& pb)
{
    SlicePtr nearest_slice = getSlice(sample->state, pb, nullptr);
    if (!nearest_slice) throw std::logic_error("[SliceBasedOracleRRT::selectTreeNode] no slices");
    if (_do_slice_ball_projection) projectSliceBall(nearest_slice->repr->state, &sample->state);
    std::vector<float> w = _distance_measure->weights();
    const uint32_t robot_mask = 1u << pb.robot_id;  // RobotStateDistanceFn
    if (active_obj_id != pb.robot_id) {
        ompl::state::SimEnvWorldState tmp(nearest_slice->repr->state);
        _oracle_sampler->sampleFeasibleState(&tmp, &sample->state, active_obj_id, _feasible_state_noise);
        _state_space->copySubState(&sample->state, &tmp, pb.robot_id);
    }
    // nearest member of the slice under the robot-only metric: slice filter on the main tree
    selected_node = _tree->nearest(sample->state, w, robot_mask, nearest_slice->id);
    ++pb.stats.num_nearest_neighbor_queries;
    if (!selected_node) selected_node = nearest_slice->repr;
}

void SliceBasedOracleRRT::addToTree(MotionPtr new_motion, MotionPtr parent, PlanningBlackboard& pb)
{
    new_motion->parent = parent;
    double slice_distance = std::numeric_limits<float>::max();
    SlicePtr closest = getSlice(new_motion->state, pb, &slice_distance);
    if (!closest || (float)slice_distance > (float)_min_slice_distance) {  // a new slice (:834)
        auto sl = std::make_shared<Slice>();
        sl->id = (int32_t)_slices_list.size();
        sl->repr = new_motion;
        sl->slice_samples_list.push_back(new_motion);
        _tree->add(new_motion, sl->id);
        auto rep = std::make_shared<ompl::planning::essentials::Motion>(_state_space->getNumObjects());
        rep->state = new_motion->state;
        _slices_nn->add(rep);
        _slices_list.push_back(sl);
    } else {
        closest->slice_samples_list.push_back(new_motion);
        _tree->add(new_motion, closest->id);
    }
}
}  // namespace algorithm

// ============================================================================== OraclePushPlanner
bool OraclePushPlanner::setup(PlanningProblem& problem)
{
    _problem = problem;
    _si = std::make_shared<algorithm::SpaceInformation>();
    _si->ctx = std::make_shared<DeviceContext>(problem.scene, problem.device);
    _si->state_space = std::make_shared<ompl::state::SimEnvWorldStateSpace>(problem.scene);
    float vnorm = std::min(problem.control_limits.velocity_limits[0], problem.control_limits.velocity_limits[1]);
    _si->control_space = std::make_shared<ompl::control::RampVelocityControlSpace>(problem.control_limits, true, vnorm);
    _si->validity_checker = std::make_shared<ompl::state::SimEnvValidityChecker>(_si->ctx);
    _si->state_propagator = std::make_shared<ompl::control::SimEnvStatePropagator>(_si->ctx);
    _si->rng = std::make_shared<util::random::RNG>(problem.seed);
    _goal = std::make_shared<ompl::state::goal::ObjectsRelocationGoal>(_si->state_space, _si->validity_checker,
                                                                       problem.relocation_goals, _si->rng);
    // createAlgorithm (:624-729)
    auto robot_oracle = std::make_shared<pushing::oracle::RampComputer>(_si->control_space);
    std::vector<pushing::oracle::ObjectData> object_data;
    for (int i = 0; i < problem.scene.n_bodies; ++i) {
        pushing::oracle::ObjectData d;
        d.mass = problem.scene.mass[i];
        d.inertia = problem.scene.inertia[i];
        d.mu = problem.scene.mu_ground[i];
        d.width = 2.0f * problem.scene.hx[i];
        d.height = problem.scene.shape[i] == MPS_SHAPE_DISC ? 2.0f * problem.scene.hx[i] : 2.0f * problem.scene.hy[i];
        object_data.push_back(d);
    }
    auto pushing_oracle = std::make_shared<pushing::oracle::HumanOracle>(robot_oracle, object_data, _si->rng);
    switch (problem.algorithm_type) {
    case AlgorithmType::Naive:
        _algorithm = std::make_shared<algorithm::NaiveRearrangementRRT>(_si, problem.num_control_samples);
        break;
    case AlgorithmType::HybridActionRRT:
        _algorithm = std::make_shared<algorithm::HybridActionRRT>(_si, pushing_oracle, robot_oracle, problem.robot_id);
        break;
    case AlgorithmType::OracleRRT:
        _algorithm = std::make_shared<algorithm::OracleRearrangementRRT>(_si, pushing_oracle, robot_oracle, problem.robot_id);
        break;
    case AlgorithmType::SliceOracleRRT: {
        auto a = std::make_shared<algorithm::SliceBasedOracleRRT>(_si, pushing_oracle, robot_oracle, problem.robot_id);
        a->setSliceBallProjection(problem.do_slice_ball_projection, pushing_oracle->getMaximalPushingDistance());
        _algorithm = a;
        break;
    }
    }
    // createShortcutAlgorithm (:731-768)
    _shortcutter.reset();
    switch (problem.shortcut_type) {
    case ShortcutType::NaiveShortcut:
        _shortcutter = std::make_shared<algorithm::NaiveShortcutter>(_si, robot_oracle);
        break;
    case ShortcutType::LocalShortcut:
        _shortcutter = std::make_shared<algorithm::LocalShortcutter>(_si, robot_oracle, problem.robot_id);
        break;
    case ShortcutType::LocalOracleShortcut:
        _shortcutter = std::make_shared<algorithm::LocalOracleShortcutter>(_si, robot_oracle, pushing_oracle, problem.robot_id);
        break;
    case ShortcutType::OracleShortcut:
        _shortcutter = std::make_shared<algorithm::OracleShortcutter>(_si, robot_oracle, pushing_oracle, problem.robot_id);
        break;
    case ShortcutType::NoShortcut:
        break;
    default:
        throw std::runtime_error("Invalid shortcut configuration encountered");
    }
    if (_shortcutter) _shortcutter->setBatchSize(problem.shortcut_batch_size);
    _is_initialized = true;
    return true;
}

bool OraclePushPlanner::solve(PlanningSolution& solution)
{
    if (!_is_initialized) throw std::logic_error("[mps::planner::pushing::OraclePushPlanner::solve] Planner not initialized. Can not plan.");
    auto pq = _algorithm->createPlanningQuery(_goal, _problem.start_state, _problem.robot_id, _problem.planning_time_out);
    pq->stopping_condition = _problem.stopping_condition;
    pq->weights = _problem.object_weights;
    pq->max_iterations = _problem.max_iterations;
    if (pq->weights.empty()) pq->weights.assign(_problem.scene.n_bodies, 1.0f);
    // problem parameters overwrite the ParamSet defaults, as strings (OraclePushPlanner.cpp:198-215)
    pq->parameters->setParam("goal_bias", f2s(_problem.goal_bias));
    pq->parameters->setParam("target_bias", f2s(_problem.target_bias));
    pq->parameters->setParam("robot_bias", f2s(_problem.robot_bias));
    pq->parameters->setParam("num_control_samples", std::to_string(_problem.num_control_samples));
    pq->parameters->setParam("action_randomness", f2s(_problem.action_noise));
    pq->parameters->setParam("state_noise", f2s(_problem.state_noise));
    solution.solved = _algorithm->plan(pq, solution.stats);
    solution.path = pq->path;
    if (solution.solved) solution.stats.reproducible = verifySolution(solution);
    _path_len_before_shortcut = solution.path ? solution.path->getNumMotions() : 0;
    if (_shortcutter && solution.solved) shortcutSolution(solution);
    return solution.solved;
}

void OraclePushPlanner::shortcutSolution(PlanningSolution& solution)
{
    // :229-238
    if (!_shortcutter || !solution.path) return;
    auto cost_fn = std::make_shared<algorithm::ActionDurationCost>(_problem.control_limits.acceleration_limits);
    algorithm::Shortcutter::ShortcutQuery sq(_goal, cost_fn, _problem.robot_id);
    sq.stopping_condition = _problem.stopping_condition;
    _shortcutter->shortcut(solution.path, sq, _problem.max_shortcut_time);
    solution.stats.cost_before_shortcut = sq.cost_before_shortcut;
    solution.stats.cost_after_shortcut = sq.cost_after_shortcut;
    solution.stats.reproducible_after_shortcut = verifySolution(solution);
}

bool OraclePushPlanner::verifySolution(PlanningSolution& solution)
{
    // replay every control through propagate and compare with the stored states; goal test on the last
    if (!solution.path || solution.path->motions.empty()) return false;
    ompl::state::SimEnvWorldState state(solution.path->motions[0]->state);
    ompl::state::SimEnvWorldState next(state.getNumObjects());
    for (size_t i = 1; i < solution.path->motions.size(); ++i) {
        float control[5];
        for (int c = 0; c < 5; ++c) control[c] = solution.path->motions[i]->control[c];
        bool ok = _si->state_propagator->propagate(&state, control, &next);
        if (!ok) return false;
        if (next.values != solution.path->motions[i]->state.values) return false;  // bit-identical replay
        state = next;
    }
    return _goal->isSatisfied(&state);
}
}  // namespace pushing
}  // namespace planner
}  // namespace mps

// ================================================================================ C entry points
using namespace mps::planner;

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

namespace {
// MPS_HOST_DEBUG=1: print a native backtrace on SIGSEGV (there is no debugger on the GPU boxes)
void segv_backtrace(int sig)
{
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "libmps_host: fatal signal, native backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
struct DebugInit {
    DebugInit()
    {
        const char* e = getenv("MPS_HOST_DEBUG");
        if (e && e[0] == '1') signal(SIGSEGV, segv_backtrace);
    }
} g_debug_init;
}  // namespace

namespace {
void problem_from_params(pushing::PlanningProblem& problem, const mps_scene* scene, const float* start,
                         const mps_goal* goals, int32_t n_goals, const mps_host_params* params)
{
    problem.scene = *scene;
    problem.start_state = ompl::state::SimEnvWorldState((unsigned int)scene->n_bodies);
    std::memcpy(problem.start_state.values.data(), start, sizeof(float) * 3 * (size_t)scene->n_bodies);
    problem.robot_id = (unsigned int)scene->robot_id;
    for (int g = 0; g < n_goals; ++g) {
        ompl::state::goal::RelocationGoalSpecification s;
        s.id = (unsigned int)goals[g].body;
        s.goal_position[0] = goals[g].x;
        s.goal_position[1] = goals[g].y;
        s.position_tolerance = goals[g].tol;
        problem.relocation_goals.push_back(s);
    }
    problem.algorithm_type = (pushing::AlgorithmType)params->algorithm;
    problem.planning_time_out = params->time_out;
    problem.goal_bias = params->goal_bias;
    problem.target_bias = params->target_bias;
    problem.robot_bias = params->robot_bias;
    problem.num_control_samples = (unsigned int)params->num_control_samples;
    problem.action_noise = params->action_randomness;
    problem.state_noise = params->state_noise;
    problem.do_slice_ball_projection = params->do_slice_ball_projection != 0;
    problem.seed = params->seed;
    problem.device = params->device;
    problem.max_iterations = (unsigned int)params->max_iterations;
    problem.shortcut_type = (pushing::ShortcutType)params->shortcut_type;
    problem.max_shortcut_time = params->max_shortcut_time;
    problem.shortcut_batch_size = params->shortcut_batch_size > 0 ? (unsigned int)params->shortcut_batch_size : 1u;
    for (int i = 0; i < 3; ++i) {
        problem.control_limits.velocity_limits[i] = params->velocity_limits[i];
        problem.control_limits.acceleration_limits[i] = scene->accel_limits[i];
    }
    problem.control_limits.duration_limits[0] = params->duration_limits[0];
    problem.control_limits.duration_limits[1] = params->duration_limits[1];
}

void fill_result(mps_host_result* result, const pushing::OraclePushPlanner& planner,
                 const pushing::PlanningSolution& solution, const mps_scene* scene, float* path_states,
                 float* path_controls, int32_t max_path)
{
    result->solved = solution.solved ? 1 : 0;
    result->reproducible = solution.stats.reproducible ? 1 : 0;
    result->num_iterations = (int32_t)solution.stats.num_iterations;
    result->num_state_propagations = (int32_t)solution.stats.num_state_propagations;
    result->num_samples = (int32_t)solution.stats.num_samples;
    result->num_nearest_neighbor_queries = (int32_t)solution.stats.num_nearest_neighbor_queries;
    result->runtime_cpu = solution.stats.runtime;
    result->runtime_wall = solution.stats.wall_time;
    result->tree_size = (int32_t)planner.getAlgorithm()->treeSize();
    auto slice_rrt = std::dynamic_pointer_cast<pushing::algorithm::SliceBasedOracleRRT>(planner.getAlgorithm());
    result->num_slices = slice_rrt ? (int32_t)slice_rrt->numSlices() : 0;
    result->cost_before_shortcut = solution.stats.cost_before_shortcut;
    result->cost_after_shortcut = solution.stats.cost_after_shortcut;
    result->reproducible_after_shortcut = solution.stats.reproducible_after_shortcut ? 1 : 0;
    if (planner.getShortcutter()) {
        const auto& st = planner.getShortcutter()->statistics();
        result->shortcut_candidates = (int32_t)st.num_candidates;
        result->shortcut_launches = (int32_t)st.num_launches;
        result->shortcut_propagations = (int32_t)st.num_propagations;
        result->shortcut_accepted = (int32_t)st.num_accepted;
    }
    result->path_len = 0;
    if (solution.solved && solution.path) {
        const size_t dim = 3 * (size_t)scene->n_bodies;
        int32_t n = (int32_t)solution.path->motions.size();
        result->path_len = n;
        for (int32_t i = 0; i < n && i < max_path; ++i) {
            if (path_states) std::memcpy(path_states + (size_t)i * dim, solution.path->motions[i]->state.values.data(), dim * sizeof(float));
            if (path_controls) std::memcpy(path_controls + (size_t)i * 5, solution.path->motions[i]->control.data(), 5 * sizeof(float));
        }
    }
}

int report(const std::exception& e, char* err, int32_t err_len)
{
    if (err && err_len > 0) {
        std::strncpy(err, e.what(), (size_t)err_len - 1);
        err[err_len - 1] = 0;
    }
    return -1;
}
}  // namespace

extern "C" int mps_host_plan(const mps_scene* scene, const float* start, const mps_goal* goals, int32_t n_goals,
                             const mps_host_params* params, mps_host_result* result, float* path_states,
                             float* path_controls, int32_t max_path, char* err, int32_t err_len)
{
    try {
        pushing::PlanningProblem problem;
        problem_from_params(problem, scene, start, goals, n_goals, params);
        pushing::OraclePushPlanner planner;
        planner.setup(problem);
        pushing::PlanningSolution solution;
        planner.solve(solution);
        std::memset(result, 0, sizeof(*result));
        fill_result(result, planner, solution, scene, path_states, path_controls, max_path);
        result->path_len_before_shortcut = (int32_t)planner.pathLengthBeforeShortcut();
        return 0;
    } catch (const std::exception& e) {
        return report(e, err, err_len);
    }
}

extern "C" int mps_host_shortcut(const mps_scene* scene, const mps_goal* goals, int32_t n_goals,
                                 const mps_host_params* params, const float* in_states, const float* in_controls,
                                 int32_t n_path, mps_host_result* result, float* path_states, float* path_controls,
                                 int32_t max_path, char* err, int32_t err_len)
{
    try {
        if (n_path < 1) throw std::invalid_argument("mps_host_shortcut: empty path");
        pushing::PlanningProblem problem;
        problem_from_params(problem, scene, in_states, goals, n_goals, params);
        pushing::OraclePushPlanner planner;
        planner.setup(problem);
        pushing::PlanningSolution solution;
        solution.path = std::make_shared<ompl::planning::essentials::Path>();
        const size_t dim = 3 * (size_t)scene->n_bodies;
        for (int32_t i = 0; i < n_path; ++i) {
            auto m = std::make_shared<ompl::planning::essentials::Motion>((unsigned int)scene->n_bodies);
            std::memcpy(m->state.values.data(), in_states + (size_t)i * dim, dim * sizeof(float));
            std::memcpy(m->control.data(), in_controls + (size_t)i * 5, 5 * sizeof(float));
            if (i > 0) m->parent = solution.path->motions.back();
            solution.path->motions.push_back(m);
        }
        solution.solved = true;
        solution.stats.reproducible = planner.verifySolution(solution);
        planner.shortcutSolution(solution);
        std::memset(result, 0, sizeof(*result));
        fill_result(result, planner, solution, scene, path_states, path_controls, max_path);
        result->path_len_before_shortcut = n_path;
        return 0;
    } catch (const std::exception& e) {
        return report(e, err, err_len);
    }
}

extern "C" void mps_host_default_params(mps_host_params* p)
{
    std::memset(p, 0, sizeof(*p));
    p->algorithm = 0;
    p->time_out = 60.0f;         // OraclePushPlanner.cpp:50
    p->goal_bias = 0.1f;         // :53-55
    p->target_bias = 0.1f;
    p->robot_bias = 0.1f;
    p->num_control_samples = 10; // :86
    p->action_randomness = 0.01f;
    p->state_noise = 0.001f;
    p->do_slice_ball_projection = 1;
    p->seed = 1;
    p->device = 0;
    p->max_iterations = 0;
    p->velocity_limits[0] = 0.6f;
    p->velocity_limits[1] = 0.6f;
    p->velocity_limits[2] = 1.5f;
    p->duration_limits[0] = 0.1f;  // :70-71
    p->duration_limits[1] = 1.0f;
    p->shortcut_type = 3;          // LocalShortcut, :87-88
    p->max_shortcut_time = 5.0f;
    p->shortcut_batch_size = 256;
}

// oracle helpers exposed for the known-answer tests of the "next" rows (RampComputer / HumanOracle)
extern "C" int mps_host_ramp_steer(const float* vel_limits, const float* accel_limits, const float* duration_limits,
                                   const float* current, const float* desired, float* controls_out, int32_t max_controls)
{
    ompl::control::RampVelocityControlSpace::ControlLimits lim;
    for (int i = 0; i < 3; ++i) {
        lim.velocity_limits[i] = vel_limits[i];
        lim.acceleration_limits[i] = accel_limits[i];
    }
    lim.duration_limits[0] = duration_limits[0];
    lim.duration_limits[1] = duration_limits[1];
    auto cs = std::make_shared<ompl::control::RampVelocityControlSpace>(lim, true, vel_limits[0]);
    pushing::oracle::RampComputer rc(cs);
    std::vector<std::array<float, 5>> out;
    rc.steer(current, desired, out);
    for (size_t i = 0; i < out.size() && (int32_t)i < max_controls; ++i)
        for (int c = 0; c < 5; ++c) controls_out[i * 5 + c] = out[i][c];
    return (int)out.size();
}

extern "C" int mps_host_project_slice_ball(const mps_scene* scene, const float* weights, int32_t robot_id, float radius,
                                           const float* center, float* sample)
{
    if (!scene || !center || !sample) return -1;
    const unsigned int B = (unsigned int)scene->n_bodies;
    auto space = std::make_shared<ompl::state::SimEnvWorldStateSpace>(*scene);
    std::vector<float> w;
    if (weights) w.assign(weights, weights + B);
    ompl::state::SimEnvWorldStateDistanceMeasure dm(space, w);
    ompl::state::SimEnvWorldState c(B), s(B);
    for (unsigned int i = 0; i < 3 * B; ++i) {
        c.values[i] = center[i];
        s.values[i] = sample[i];
    }
    const bool inside = pushing::algorithm::SliceBasedOracleRRT::projectSliceBall(dm, B, (unsigned int)robot_id, radius, c, &s);
    for (unsigned int i = 0; i < 3 * B; ++i) sample[i] = s.values[i];
    return inside ? 1 : 0;
}

extern "C" float mps_host_max_pushing_distance(void)
{
    auto cs = std::make_shared<ompl::control::RampVelocityControlSpace>(ompl::control::RampVelocityControlSpace::ControlLimits(), true, 0.6f);
    auto rc = std::make_shared<pushing::oracle::RampComputer>(cs);
    pushing::oracle::HumanOracle h(rc, {}, std::make_shared<util::random::RNG>(1));
    return h.getMaximalPushingDistance();
}
