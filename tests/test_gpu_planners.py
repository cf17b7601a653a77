"""The four planners of the reference (pushing/algorithm/RRT.cpp) driven through the host-side C++ mirror
on the device: they run, their statistics are consistent, and every solution replays bit-identically —
both through the device propagator (verifySolution, OraclePushPlanner.cpp:495-519) and through the CPU
oracle."""
import numpy as np
import pytest

from manipulation_planning_suite_b200 import host_api, scenes

import oracle_binding as ob

pytestmark = pytest.mark.gpu


def _problem(seed=1001):
    sc, start = scenes.make_scene(3, seed=seed)
    # an easy relocation: object 1 has to move 0.3 m
    goal = (1, float(start[3]) + 0.25, float(start[4]) + 0.15, 0.1)
    return sc, start, [goal]


def _check_solution(sc, goals, r):
    assert r.reproducible, "device replay of the path is not bit-identical"
    st = r.path_states[0]
    for i in range(1, len(r.path_states)):
        ok, out, c, fl, _ = ob.propagate(sc, st, r.path_controls[i])
        assert ok
        assert np.array_equal(out, r.path_states[i]), f"oracle replay differs at motion {i}"
        assert c[4] == r.path_controls[i][4]
        st = out
    g = goals[0]
    assert np.hypot(st[3 * g[0]] - g[1], st[3 * g[0] + 1] - g[2]) <= g[3] + 1e-6


@pytest.mark.parametrize("algorithm", ["Naive", "HybridActionRRT", "OracleRRT", "SliceOracleRRT"])
def test_planner_runs_and_solutions_replay(algorithm):
    sc, start, goals = _problem()
    solved_any = False
    for seed in (1, 2, 3):
        # NaiveRRT only has random controls: it needs far more iterations than the oracle-guided planners
        iters = 6000 if algorithm == "Naive" else 400
        r = host_api.plan(sc, start, goals, algorithm=algorithm, seed=seed, max_iterations=iters, time_out=60.0,
                          goal_bias=0.2, target_bias=0.3, robot_bias=0.1, num_control_samples=10)
        s = r.stats
        assert s["num_iterations"] >= 1 and s["num_samples"] == s["num_iterations"]
        assert s["tree_size"] >= 1
        if algorithm == "Naive":
            assert s["num_state_propagations"] == 10 * s["num_iterations"]
        if algorithm == "SliceOracleRRT":
            assert s["num_slices"] >= 1
        if r.solved:
            solved_any = True
            assert len(r.path_states) >= 2
            assert np.array_equal(r.path_states[0], start)
            _check_solution(sc, goals, r)
    assert solved_any, f"{algorithm} solved none of 3 seeds of an easy problem"


def test_start_in_goal_is_solved_without_extending():
    sc, start, _ = _problem()
    goals = [(1, float(start[3]), float(start[4]), 0.1)]
    r = host_api.plan(sc, start, goals, algorithm="Naive", max_iterations=10)
    assert r.solved and r.stats["num_iterations"] == 0 and len(r.path_states) == 1


def test_large_k_naive_rrt():
    """num_control_samples = 1024: the K-loop of NaiveControlSampler::sampleTo as one launch."""
    sc, start, goals = _problem()
    r = host_api.plan(sc, start, goals, algorithm="Naive", seed=5, max_iterations=40, num_control_samples=1024,
                      goal_bias=0.3, target_bias=0.5)
    assert r.stats["num_state_propagations"] == 1024 * r.stats["num_iterations"]
    if r.solved:
        _check_solution(sc, goals, r)


# ---------------------------------------------------------------------------------------------------
# shortcutters (pushing/algorithm/RRT.cpp:889-1688): candidate rounds as batched replays
# ---------------------------------------------------------------------------------------------------
ACCEL = np.array([2.0, 2.0, 4.0], np.float32)


def _duration(c):
    """RampVelocityControl::getMaxDuration (RampVelocityControl.cpp:112-115) in float arithmetic."""
    c = np.asarray(c, np.float32)
    t_acc = np.float32(np.max(np.abs(c[:3] / ACCEL)))
    return np.float32(np.float32(np.float32(2.0) * t_acc + c[3]) + c[4])


def _path_cost(controls, limit=-1):
    """CostFunction::cost(path, limit) with ActionDurationCost (Essentials.cpp:197-211, Costs.cpp:12-17)."""
    n = len(controls)
    if limit >= 0:
        n = min(limit + 1, n)
    if n <= 1:
        return 0.0
    return float(sum(float(_duration(controls[i])) for i in range(n)))


def _goal_satisfied(state, goals):
    d = 0.0
    for (b, x, y, tol) in goals:
        fx = np.float32(x) - state[3 * b]
        fy = np.float32(y) - state[3 * b + 1]
        d += max(float(np.sqrt(float(fx) ** 2 + float(fy) ** 2)) - float(np.float32(tol)), 0.0)
    return d <= np.finfo(np.float64).eps


def _local_shortcut_sequential(sc, goals, states, controls, robot=0):
    """LocalShortcutter::shortcut (RRT.cpp:1189-1271) one candidate at a time on the CPU oracle."""
    states = [s.copy() for s in states]
    controls = [c.copy() for c in controls]
    if len(states) <= 2:
        return states, controls, None, None
    cost = _path_cost(controls)
    initial = cost
    stride = 2
    finished = False
    while not finished:
        sid = 0
        while sid + stride < len(states):
            eid = sid + stride
            new = host_api.ramp_steer(states[sid][3 * robot:3 * robot + 3], states[eid][3 * robot:3 * robot + 3])
            queue = [c for c in new] + [c.copy() for c in controls[eid + 1:]]
            t_states, t_controls = [states[sid]], [controls[sid]]
            st, valid, goal = states[sid], True, False
            for c in queue:
                ok, out, cc, _, _ = ob.propagate(sc, st, c)
                if not ok:
                    valid = False
                    break
                t_states.append(out)
                t_controls.append(cc)
                st = out
                if _goal_satisfied(out, goals):
                    goal = True
                    break
            new_cost = _path_cost(controls, sid) + _path_cost(t_controls)
            if valid and goal and new_cost < cost:
                states = states[:sid] + t_states
                controls = controls[:sid] + t_controls
                cost = new_cost
            sid += 1
        stride += 1
        finished = stride >= len(states)
    return states, controls, initial, cost


def _solved_path(algorithm="OracleRRT"):
    sc, start, goals = _problem()
    for seed in range(1, 12):
        r = host_api.plan(sc, start, goals, algorithm=algorithm, seed=seed, max_iterations=600, goal_bias=0.2,
                          target_bias=0.3, robot_bias=0.1)
        if r.solved and len(r.path_states) >= 5:
            return sc, goals, r
    pytest.skip("no seed gave a solution with at least 5 motions")


def test_local_shortcutter_batched_equals_sequential_reference_loop():
    sc, goals, r = _solved_path()
    got = host_api.shortcut(sc, goals, r.path_states, r.path_controls, "LocalShortcut", max_shortcut_time=600.0)
    one = host_api.shortcut(sc, goals, r.path_states, r.path_controls, "LocalShortcut", max_shortcut_time=600.0,
                            shortcut_batch_size=1)
    assert np.array_equal(got.path_states, one.path_states) and np.array_equal(got.path_controls, one.path_controls)
    assert got.stats["shortcut_launches"] < one.stats["shortcut_launches"]
    ws, wc, initial, cost = _local_shortcut_sequential(sc, goals, list(r.path_states), list(r.path_controls))
    assert np.array_equal(got.path_states, np.array(ws)), "batched shortcut differs from the sequential loop"
    assert np.array_equal(got.path_controls, np.array(wc))
    assert got.stats["cost_before_shortcut"] == np.float32(initial)
    assert got.stats["cost_after_shortcut"] == np.float32(cost)
    assert got.stats["cost_after_shortcut"] <= got.stats["cost_before_shortcut"]
    assert got.stats["reproducible_after_shortcut"] == 1
    _check_solution(sc, goals, got)


@pytest.mark.parametrize("kind", ["NaiveShortcut", "LocalOracleShortcut", "OracleShortcut"])
def test_shortcutters_keep_the_goal_and_never_raise_the_cost(kind):
    sc, goals, r = _solved_path("SliceOracleRRT")
    got = host_api.shortcut(sc, goals, r.path_states, r.path_controls, kind, max_shortcut_time=600.0, seed=3)
    s = got.stats
    assert s["reproducible_after_shortcut"] == 1
    assert s["cost_after_shortcut"] <= s["cost_before_shortcut"]
    assert s["cost_before_shortcut"] == np.float32(_path_cost(r.path_controls))
    assert s["cost_after_shortcut"] == np.float32(_path_cost(got.path_controls))
    assert s["shortcut_candidates"] >= 1 and s["shortcut_launches"] >= 1
    assert np.array_equal(got.path_states[0], r.path_states[0])
    _check_solution(sc, goals, got)
    if kind == "NaiveShortcut":  # the pair order is drawn from the planner's generator: batch size must not matter
        one = host_api.shortcut(sc, goals, r.path_states, r.path_controls, kind, max_shortcut_time=600.0, seed=3,
                                shortcut_batch_size=1)
        assert np.array_equal(got.path_states, one.path_states)
        assert np.array_equal(got.path_controls, one.path_controls)


def test_plan_with_reference_default_shortcut():
    """OraclePushPlanner::solve with shortcut_type = LocalShortcut (the reference default, OraclePushPlanner.cpp:87)."""
    sc, start, goals = _problem()
    for seed in range(1, 8):
        r = host_api.plan(sc, start, goals, algorithm="OracleRRT", seed=seed, max_iterations=600, goal_bias=0.2,
                          target_bias=0.3, robot_bias=0.1, shortcut_type="LocalShortcut")
        if r.solved:
            s = r.stats
            assert s["reproducible"] == 1 and s["reproducible_after_shortcut"] == 1
            assert s["path_len"] <= s["path_len_before_shortcut"]
            assert s["cost_after_shortcut"] <= s["cost_before_shortcut"]
            _check_solution_states_only(sc, goals, r)
            return
    pytest.fail("no seed solved the easy problem")


def _check_solution_states_only(sc, goals, r):
    st = r.path_states[0]
    for i in range(1, len(r.path_states)):
        ok, out, _, _, _ = ob.propagate(sc, st, r.path_controls[i])
        assert ok and np.array_equal(out, r.path_states[i])
        st = out
    assert _goal_satisfied(st, goals)


# ---------------------------------------------------------------------------------------------------
# DataGenerator (pushing/oracle/DataGenerator.cpp:56-116): one batched launch per file
# ---------------------------------------------------------------------------------------------------
def test_data_generator_triplets_replay_on_the_oracle(tmp_path):
    from manipulation_planning_suite_b200 import data_generator as dg
    sc, _ = scenes.make_scene(1, seed=4)          # robot + one object, as the reference's generator worlds
    gen = dg.DataGenerator(sc, object_id=1, seed=11)
    f = tmp_path / "train.csv"
    info = gen.generate_data(str(f), 96, annotation="0.35, 0.19")
    assert info["propagations"] == 96 and 0 < info["lines"] <= 96
    assert info["launches"] <= 30, "state validation and propagation must be batched (rejection rounds + 1 extend)"
    s, r, c, notes = dg.read_triplets(str(f), sc.n_bodies)
    assert len(s) == info["lines"] and notes[0] == "0.35, 0.19"
    moved = 0
    for i in range(len(s)):
        assert np.all(s[i, 0:3] == 0.0)                                       # robot at the origin (:203-206)
        assert np.hypot(s[i, 3], s[i, 4]) < gen.max_distance                  # object within reach (:216-217)
        ok, out, cc, _, _ = ob.propagate(sc, s[i], c[i])
        assert ok and np.array_equal(out, r[i]) and cc[4] == c[i, 4]
        moved += int(np.abs(r[i, 3:6] - s[i, 3:6]).max() > 1e-5)
    assert moved >= 3, "no sampled push moved the object: weak test"
    # state + dynamics noise: per-sample worlds, same file format
    noisy = dg.DataGenerator(sc, 1, dg.Parameters(0.005, 0.05, 0.05, 0.02), seed=12)
    info2 = noisy.generate_data(str(f), 8, deterministic=False, num_noise_samples=3)
    s2, r2, c2, _ = dg.read_triplets(str(f), sc.n_bodies)
    assert info2["propagations"] <= 24 and len(s2) == info2["lines"] > 0
    gen.close()
    noisy.close()
