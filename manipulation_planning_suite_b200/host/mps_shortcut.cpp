// mps_shortcut.cpp — the reference's path shortcutters (pushing/algorithm/RRT.cpp:889-1688) with the
// candidate evaluation moved onto the device.
//
// Reference control flow, per candidate (start_id, end_id) of a solution path:
//   1. new actions that move the robot (and, for the oracle variants, one object) from the state at start_id
//      towards the state at end_id are computed on the host (RampComputer / HumanOracle),
//   2. forwardPropagatePath (:946-1034) propagates them and then the old path's actions from end_id + 1 on,
//      stopping at the first failed propagation or the first goal state,
//   3. the candidate is accepted when the goal was reached at lower ActionDurationCost.
// Candidates of one round are independent of each other until one is accepted, so they are propagated together:
// every candidate is one control chain with its own start state (mps_extend_in.starts, stop_at_goal) and the
// results are then inspected in the reference's order.  The candidate accepted is the one the reference's
// sequential loop accepts first; work done for the candidates behind it is discarded.
#include <algorithm>
#include <cstring>

#include "mps_host.hpp"

namespace mps {
namespace planner {
namespace pushing {
namespace algorithm {
using ompl::planning::essentials::Motion;
using ompl::planning::essentials::MotionPtr;
using ompl::planning::essentials::Path;
using ompl::planning::essentials::PathPtr;
using ompl::state::SimEnvWorldState;

// ---- costs -------------------------------------------------------------------------------------
double CostFunction::cost(const Path& path, int limit) const
{
    double value = 0.0;
    unsigned int num_motions = path.getNumMotions();
    if (limit >= 0) num_motions = std::min((unsigned int)limit + 1, num_motions);
    if (num_motions <= 1) return 0.0;
    const Motion* prev = path.motions[0].get();
    for (unsigned int i = 0; i < num_motions; ++i) {
        const Motion* current = path.motions[i].get();
        value += cost(*prev, *current);
        prev = current;
    }
    return value;
}

ActionDurationCost::ActionDurationCost(const float accel_limits[3])
{
    for (int i = 0; i < 3; ++i) _accel_limits[i] = accel_limits[i];
}

double ActionDurationCost::duration(const std::array<float, 5>& control) const
{
    ompl::control::RampVelocityControl ramp(_accel_limits);
    ramp.setParameters(control.data());
    return (double)ramp.getMaxDuration();
}

double ActionDurationCost::cost(const Motion&, const Motion& second) const
{
    return duration(second.control);
}

// ---- Shortcutter -------------------------------------------------------------------------------
Shortcutter::Shortcutter(SpaceInformationPtr si) : _si(si) {}

void Shortcutter::initCandidate(Candidate& c, const Path& current_path, unsigned int start_id, unsigned int end_id) const
{
    c.start_id = start_id;
    c.end_id = end_id;
    c.object_id = 0;
    c.stage = 0;
    c.pending.clear();
    c.appended.clear();
    c.last = current_path.motions[start_id]->state;
    c.valid = true;
    c.goal = false;
    c.done = false;
}

void Shortcutter::propagateCandidates(std::vector<Candidate>& cands, const Path& current_path, ShortcutQuery& sq)
{
    const unsigned int n_objects = _si->state_space->getNumObjects();
    const size_t dim = 3 * (size_t)n_objects;
    std::vector<size_t> live;
    std::vector<float> starts, controls, all_states, all_rest;
    std::vector<int32_t> n_ctrl;
    std::vector<uint32_t> all_flags;
    _stats.num_candidates += (unsigned int)cands.size();
    for (;;) {
        live.clear();
        size_t L = 0;
        for (size_t i = 0; i < cands.size(); ++i) {
            Candidate& c = cands[i];
            // ask for the next stage until there is something to propagate (a stage may be empty)
            while (!c.done && c.pending.empty()) {
                bool more = nextControls(c, current_path, sq);
                if (!c.valid || !more) c.done = true;
            }
            if (c.done) continue;
            live.push_back(i);
            L = std::max(L, c.pending.size());
        }
        if (live.empty()) break;
        L = std::min(L, (size_t)MPS_MAX_CHAIN);
        const size_t K = live.size();
        starts.assign(K * dim, 0.0f);
        controls.assign(K * L * 5, 0.0f);
        n_ctrl.assign(K, 0);
        all_states.resize(K * L * dim);
        all_rest.resize(K * L);
        all_flags.resize(K * L);
        for (size_t k = 0; k < K; ++k) {
            const Candidate& c = cands[live[k]];
            std::memcpy(starts.data() + k * dim, c.last.values.data(), dim * sizeof(float));
            size_t n = std::min(c.pending.size(), L);
            n_ctrl[k] = (int32_t)n;
            for (size_t l = 0; l < n; ++l)
                std::memcpy(controls.data() + (k * L + l) * 5, c.pending[l].data(), 5 * sizeof(float));
        }
        mps_extend_in in;
        std::memset(&in, 0, sizeof(in));
        in.starts = starts.data();
        in.controls = controls.data();
        in.n_ctrl = n_ctrl.data();
        in.K = (int32_t)K;
        in.L = (int32_t)L;
        in.mask = 0xffffffffu;
        in.goals = sq.goal_region->abiGoals().data();
        in.n_goals = (int32_t)sq.goal_region->abiGoals().size();
        in.stop_at_goal = 1;
        mps_extend_out out;
        std::memset(&out, 0, sizeof(out));
        out.all_states = all_states.data();
        out.all_rest = all_rest.data();
        out.all_flags = all_flags.data();
        _si->ctx->check(mps_extend(_si->ctx->get(), &in, &out), "Shortcutter::propagateCandidates");
        _stats.num_launches++;
        for (size_t k = 0; k < K; ++k) {
            Candidate& c = cands[live[k]];
            for (int32_t l = 0; l < n_ctrl[k]; ++l) {
                const uint32_t f = all_flags[k * L + l];
                if (!(f & MPS_F_RAN)) break;
                _stats.num_propagations++;
                if (!(f & MPS_F_OK)) {  // propagation failed: the candidate is dead (RRT.cpp:961-963, :1014-1016)
                    c.valid = false;
                    c.done = true;
                    break;
                }
                MotionPtr m = std::make_shared<Motion>(n_objects);
                std::memcpy(m->state.values.data(), all_states.data() + (k * L + l) * dim, dim * sizeof(float));
                m->control = c.pending.front();
                m->control[4] = all_rest[k * L + l];  // rest time written back by propagate
                c.pending.pop_front();
                c.appended.push_back(m);
                c.last = m->state;
                if (f & MPS_F_GOAL) {  // :966-968, :1011-1013
                    c.goal = true;
                    c.done = true;
                    break;
                }
            }
        }
    }
}

void Shortcutter::finish(PathPtr iopath, const Path& current_path, ShortcutQuery& sq, double initial_cost,
                         double cost) const
{
    iopath->motions = current_path.motions;
    for (size_t i = 0; i < iopath->motions.size(); ++i)
        iopath->motions[i]->parent = i > 0 ? iopath->motions[i - 1] : nullptr;
    sq.cost_before_shortcut = (float)initial_cost;
    sq.cost_after_shortcut = (float)cost;
}

namespace {
// Path::deepCopy (Essentials.cpp:183-194)
Path deep_copy(const Path& p)
{
    Path out;
    for (const auto& m : p.motions) {
        MotionPtr n = std::make_shared<Motion>(*m);
        n->parent = out.motions.empty() ? nullptr : out.motions.back();
        out.motions.push_back(n);
    }
    return out;
}

// robot actions of the plain shortcutters (computeRobotActions :1136-1153, :1289-1305) followed by the rest of
// the old path (:987-1019)
bool robot_steer_stages(const oracle::RampComputer& steerer, unsigned int robot_id,
                        std::deque<std::array<float, 5>>& pending, int& stage, const SimEnvWorldState& last,
                        unsigned int end_id, const Path& current_path)
{
    if (stage == 0) {
        std::vector<std::array<float, 5>> params;
        steerer.steer(last.getObjectState(robot_id), current_path.motions[end_id]->state.getObjectState(robot_id), params);
        pending.assign(params.begin(), params.end());
        stage = 1;
        return true;
    }
    if (stage == 1) {
        for (unsigned int i = end_id + 1; i < current_path.getNumMotions(); ++i)
            pending.push_back(current_path.motions[i]->control);
        stage = 2;
        return true;
    }
    return false;
}
}  // namespace

// ---- NaiveShortcutter --------------------------------------------------------------------------
NaiveShortcutter::NaiveShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle)
    : Shortcutter(si), _robot_oracle(robot_oracle)
{
}

bool NaiveShortcutter::nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq)
{
    return robot_steer_stages(*_robot_oracle, sq.robot_id, c.pending, c.stage, c.last, c.end_id, current_path);
}

void NaiveShortcutter::createAllPairs(std::vector<std::pair<unsigned int, unsigned int>>& all_pairs, unsigned int n) const
{
    all_pairs.clear();
    if (n <= 2) return;
    all_pairs.reserve((n * n - 3 * n + 2) / 2);
    for (unsigned int i = 0; i < n - 2; ++i)
        for (unsigned int j = i + 2; j < n; ++j) all_pairs.push_back(std::make_pair(i, j));
    // the reference seeds a fresh mt19937 from random_device (:1168-1170); here the planner's generator seeds it
    std::mt19937 g((uint32_t)_si->rng->uniformInt(0, 0x7fffffff));
    std::shuffle(all_pairs.begin(), all_pairs.end(), g);
}

void NaiveShortcutter::shortcut(PathPtr iopath, ShortcutQuery& sq, float max_time)
{
    std::vector<std::pair<unsigned int, unsigned int>> all_pairs;
    bool finished = iopath->getNumMotions() <= 2;
    if (finished) return;
    Path current_path = deep_copy(*iopath);
    double current_path_cost = sq.cost_function->cost(current_path);
    const double initial_cost = current_path_cost;
    _timer.startTimer(max_time);
    std::vector<Candidate> cands;
    while (!_timer.timeOutExceeded() && !finished && !sq.stopping_condition()) {
        createAllPairs(all_pairs, current_path.getNumMotions());
        size_t pos = 0;
        bool accepted = false;
        while (!_timer.timeOutExceeded() && pos < all_pairs.size() && !sq.stopping_condition() && !accepted) {
            const size_t n_batch = std::min((size_t)_batch_size, all_pairs.size() - pos);
            cands.assign(n_batch, Candidate());
            for (size_t b = 0; b < n_batch; ++b)
                initCandidate(cands[b], current_path, all_pairs[pos + b].first, all_pairs[pos + b].second);
            propagateCandidates(cands, current_path, sq);
            for (size_t b = 0; b < n_batch; ++b) {
                const Candidate& c = cands[b];
                // new_path: all waypoints up to start_id (incl) + whatever was appended (:1096-1102)
                Path new_path;
                new_path.motions.assign(current_path.motions.begin(), current_path.motions.begin() + c.start_id + 1);
                new_path.motions.insert(new_path.motions.end(), c.appended.begin(), c.appended.end());
                const bool successful_path = c.goal;  // second member of forwardPropagatePath's result
                const double new_cost = sq.cost_function->cost(new_path);
                if (successful_path && new_cost < current_path_cost) {
                    current_path = new_path;
                    current_path_cost = new_cost;
                    finished = current_path.getNumMotions() <= 2;
                    accepted = true;
                    _stats.num_accepted++;
                    break;
                }
            }
            pos += n_batch;
        }
        // finished if all pairs were tried without success (:1118)
        finished |= (!accepted && pos >= all_pairs.size());
    }
    finish(iopath, current_path, sq, initial_cost, current_path_cost);
}

// ---- LocalShortcutter --------------------------------------------------------------------------
LocalShortcutter::LocalShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle, unsigned int robot_id)
    : Shortcutter(si), _robot_oracle(robot_oracle), _robot_id(robot_id)
{
}

bool LocalShortcutter::nextControls(Candidate& c, const Path& current_path, ShortcutQuery&)
{
    return robot_steer_stages(*_robot_oracle, _robot_id, c.pending, c.stage, c.last, c.end_id, current_path);
}

void LocalShortcutter::shortcut(PathPtr iopath, ShortcutQuery& sq, float max_time)
{
    bool finished = iopath->getNumMotions() <= 2;
    if (finished) return;
    Path current_path = deep_copy(*iopath);
    double current_path_cost = sq.cost_function->cost(current_path);
    const double initial_cost = current_path_cost;
    unsigned int stride = 2;
    _timer.startTimer(max_time);
    std::vector<Candidate> cands;
    while (!_timer.timeOutExceeded() && !finished && !sq.stopping_condition()) {
        unsigned int start_id = 0;
        while (start_id + stride < current_path.getNumMotions() && !_timer.timeOutExceeded() && !sq.stopping_condition()) {
            // all start ids of this stride that are still ahead are evaluated against the current path; the
            // results behind an accepted one are stale (the path changed under them) and are evaluated again
            const unsigned int n = current_path.getNumMotions();
            const size_t n_batch = std::min((size_t)_batch_size, (size_t)(n - stride - start_id));
            cands.assign(n_batch, Candidate());
            for (size_t b = 0; b < n_batch; ++b)
                initCandidate(cands[b], current_path, start_id + (unsigned int)b, start_id + (unsigned int)b + stride);
            propagateCandidates(cands, current_path, sq);
            bool accepted = false;
            for (size_t b = 0; b < n_batch && !accepted; ++b) {
                const Candidate& c = cands[b];
                // trailing_path = first_wp + appended motions (:1223-1231)
                Path trailing_path;
                trailing_path.motions.push_back(current_path.motions[c.start_id]);
                trailing_path.motions.insert(trailing_path.motions.end(), c.appended.begin(), c.appended.end());
                const bool successful_path = c.valid && c.goal;
                // :1238 — the action stored in first_wp is counted by both terms, as in the reference
                const double new_cost = sq.cost_function->cost(current_path, (int)c.start_id) + sq.cost_function->cost(trailing_path);
                if (successful_path && new_cost < current_path_cost) {
                    Path new_path;
                    new_path.motions.assign(current_path.motions.begin(), current_path.motions.begin() + c.start_id);
                    new_path.motions.insert(new_path.motions.end(), trailing_path.motions.begin(), trailing_path.motions.end());
                    current_path = new_path;
                    current_path_cost = new_cost;
                    finished = current_path.getNumMotions() <= 2;
                    accepted = true;
                    _stats.num_accepted++;
                }
                start_id = c.start_id + 1;
            }
        }
        ++stride;
        finished = stride >= current_path.getNumMotions();
    }
    finish(iopath, current_path, sq, initial_cost, current_path_cost);
}

// ---- LocalOracleShortcutter --------------------------------------------------------------------
namespace {
// distance of one object's sub-space (SimEnvObjectStateSpace): the world measure with only that object active
float object_distance(const mps_scene& scene, const SimEnvWorldState& a, const SimEnvWorldState& b, unsigned int i)
{
    double d = 0.0;
    std::vector<float> w((size_t)scene.n_bodies, 1.0f);
    mps_distance(&scene, a.values.data(), b.values.data(), w.data(), 1u << i, &d);
    return (float)d;
}

// stages shared by the two oracle shortcutters (computeShortcut :1347-1411 and :1573-1649):
//   0 steer the robot to a state from which `object_id` can be pushed towards its pose in the end state,
//   1 the oracle's push, 2 steer the robot to its pose in the end state; with object_id == robot only 2.
// Returns false when the stage cannot produce the controls the reference requires.
bool oracle_stage(oracle::OracleControlSampler& sampler, unsigned int robot_id, unsigned int object_id, int stage,
                  const SimEnvWorldState& last, const SimEnvWorldState& end_state,
                  std::deque<std::array<float, 5>>& pending, bool& required_empty)
{
    std::vector<std::array<float, 5>> controls;
    required_empty = false;
    if (stage == 0) {
        if (object_id == robot_id) return true;
        SimEnvWorldState feasible(last);
        sampler.sampleFeasibleState(&feasible, &end_state, object_id);
        sampler.steerRobot(controls, &last, &feasible);
        required_empty = controls.empty();  // "OracleControlSampler provided no controls at all"
    } else if (stage == 1) {
        if (object_id == robot_id) return true;
        sampler.steerPush(controls, &last, &end_state, object_id);
    } else {
        sampler.steerRobot(controls, &last, &end_state);
    }
    pending.assign(controls.begin(), controls.end());
    return true;
}
}  // namespace

LocalOracleShortcutter::LocalOracleShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle,
                                               oracle::PushingOraclePtr pushing_oracle, unsigned int robot_id)
    : LocalShortcutter(si, robot_oracle, robot_id)
{
    _oracle_sampler = std::make_shared<oracle::OracleControlSampler>(si->state_space, si->control_space, pushing_oracle,
                                                                     robot_oracle, robot_id, si->rng);
}

unsigned int LocalOracleShortcutter::selectObject(const SimEnvWorldState& start_state, const SimEnvWorldState& end_state) const
{
    unsigned int object_id = _robot_id;
    float max_distance = 0.0f;
    for (unsigned int i = 0; i < _si->state_space->getNumObjects(); ++i) {
        if (i == _robot_id) continue;
        float dist = object_distance(_si->state_space->scene(), start_state, end_state, i);
        if (dist > max_distance) {
            max_distance = dist;
            object_id = i;
        }
    }
    return object_id;
}

bool LocalOracleShortcutter::nextControls(Candidate& c, const Path& current_path, ShortcutQuery&)
{
    const SimEnvWorldState& end_state = current_path.motions[c.end_id]->state;
    if (c.stage == 0) c.object_id = selectObject(c.last, end_state);  // c.last is the state at start_id here
    if (c.stage <= 2) {
        bool required_empty = false;
        oracle_stage(*_oracle_sampler, _robot_id, c.object_id, c.stage, c.last, end_state, c.pending, required_empty);
        if (required_empty) c.valid = false;  // :1367-1370
        c.stage++;
        return true;
    }
    if (c.stage == 3) {  // forwardPropagatePath(trailing_path, current_path, end_id + 1) :1229-1231
        for (unsigned int i = c.end_id + 1; i < current_path.getNumMotions(); ++i)
            c.pending.push_back(current_path.motions[i]->control);
        c.stage = 4;
        return true;
    }
    return false;
}

// ---- OracleShortcutter -------------------------------------------------------------------------
OracleShortcutter::OracleShortcutter(SpaceInformationPtr si, oracle::RobotOraclePtr robot_oracle,
                                     oracle::PushingOraclePtr pushing_oracle, unsigned int robot_id)
    : Shortcutter(si)
{
    _oracle_sampler = std::make_shared<oracle::OracleControlSampler>(si->state_space, si->control_space, pushing_oracle,
                                                                     robot_oracle, robot_id, si->rng);
}

void OracleShortcutter::fillPairQueue(PairQueue& all_pairs, unsigned int n) const
{
    all_pairs.clear();
    if (n <= 2) return;
    for (unsigned int i = 0; i < n - 2; ++i)
        for (unsigned int j = i + 2; j < n; ++j) all_pairs.push_back(std::make_pair(i, j));
    std::mt19937 g((uint32_t)_si->rng->uniformInt(0, 0x7fffffff));  // random_device in the reference (:1516-1518)
    std::shuffle(all_pairs.begin(), all_pairs.end(), g);
}

unsigned int OracleShortcutter::selectObject(const Path& current_path, const IndexPair& current_pair,
                                             PairMap& pair_to_objects, PairQueue& pair_queue, unsigned int robot_id) const
{
    auto map_iter = pair_to_objects.find(current_pair);
    ObjectIds object_ids;
    if (map_iter == pair_to_objects.end()) {
        // first time this pair is tried: objects that differ between the two states, largest distance last
        const SimEnvWorldState& first_state = current_path.motions[current_pair.first]->state;
        const SimEnvWorldState& snd_state = current_path.motions[current_pair.second]->state;
        std::vector<std::pair<unsigned int, float>> distances;
        for (unsigned int i = 0; i < _si->state_space->getNumObjects(); ++i) {
            if (i == robot_id) continue;
            float dist = object_distance(_si->state_space->scene(), first_state, snd_state, i);
            if (dist > 0.0f) distances.push_back(std::make_pair(i, dist));
        }
        std::sort(distances.begin(), distances.end(),
                  [](std::pair<unsigned int, float> a, std::pair<unsigned int, float> b) { return a.second < b.second; });
        for (auto& d : distances) object_ids.push_back(d.first);
    } else {
        object_ids = map_iter->second;
    }
    if (object_ids.empty()) return robot_id;  // only move the robot
    unsigned int target_id = object_ids.back();
    object_ids.pop_back();
    pair_to_objects[current_pair] = object_ids;
    if (!object_ids.empty()) pair_queue.push_back(current_pair);  // try this pair again with the next object
    return target_id;
}

bool OracleShortcutter::nextControls(Candidate& c, const Path& current_path, ShortcutQuery& sq)
{
    const SimEnvWorldState& end_state = current_path.motions[c.end_id]->state;
    if (c.stage <= 2) {
        bool required_empty = false;
        oracle_stage(*_oracle_sampler, sq.robot_id, c.object_id, c.stage, c.last, end_state, c.pending, required_empty);
        // extendPath with no controls reports failure (:1658, :1687), so an empty robot steer ends the candidate
        if (required_empty || (c.stage == 2 && c.pending.empty())) c.valid = false;
        c.stage++;
        return true;
    }
    if (c.stage == 3) {  // remaining path (:1643-1648); nothing left to propagate counts as failure as well
        for (unsigned int i = c.end_id + 1; i < current_path.getNumMotions(); ++i)
            c.pending.push_back(current_path.motions[i]->control);
        if (c.pending.empty()) c.valid = false;
        c.stage = 4;
        return true;
    }
    return false;
}

void OracleShortcutter::shortcut(PathPtr iopath, ShortcutQuery& sq, float max_time)
{
    PairQueue all_pairs;
    PairMap pair_to_objects;
    bool finished = iopath->getNumMotions() <= 2;
    if (finished) return;
    Path current_path = deep_copy(*iopath);
    double current_path_cost = sq.cost_function->cost(current_path);
    const double initial_cost = current_path_cost;
    const unsigned int robot_id = sq.robot_id;
    _timer.startTimer(max_time);
    std::vector<Candidate> cands;
    while (!_timer.timeOutExceeded() && !finished && !sq.stopping_condition()) {
        fillPairQueue(all_pairs, current_path.getNumMotions());
        pair_to_objects.clear();
        bool accepted = false;
        while (!_timer.timeOutExceeded() && !all_pairs.empty() && !sq.stopping_condition() && !accepted) {
            // pop the next candidates in queue order; selectObject may queue a pair again for its next object
            cands.clear();
            while (!all_pairs.empty() && cands.size() < _batch_size) {
                IndexPair current_pair = all_pairs.front();
                all_pairs.pop_front();
                unsigned int object_id = selectObject(current_path, current_pair, pair_to_objects, all_pairs, robot_id);
                cands.emplace_back();
                initCandidate(cands.back(), current_path, current_pair.first, current_pair.second);
                cands.back().object_id = object_id;
            }
            propagateCandidates(cands, current_path, sq);
            for (const Candidate& c : cands) {
                Path new_path;
                new_path.motions.assign(current_path.motions.begin(), current_path.motions.begin() + c.start_id + 1);
                new_path.motions.insert(new_path.motions.end(), c.appended.begin(), c.appended.end());
                const bool successful_path = c.valid && c.goal;
                const double new_cost = sq.cost_function->cost(new_path);
                if (successful_path && new_cost < current_path_cost) {
                    current_path = new_path;
                    current_path_cost = new_cost;
                    finished = current_path.getNumMotions() <= 2;
                    accepted = true;
                    _stats.num_accepted++;
                    break;
                }
            }
        }
        finished |= (!accepted && all_pairs.empty());
    }
    finish(iopath, current_path, sq, initial_cost, current_path_cost);
}
}  // namespace algorithm
}  // namespace pushing
}  // namespace planner
}  // namespace mps
