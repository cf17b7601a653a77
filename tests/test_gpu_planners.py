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
