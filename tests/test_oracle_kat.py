"""CPU tests: the oracle against the golden known-answer vectors (tests/golden/kat_vectors.json, made by
tests/golden/make_kat.py from the reference's in-tree formulas) and against hand-derived constants."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from manipulation_planning_suite_b200 import abi, scenes

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_vectors.json")))


def _ramp(params, alim):
    r = ob.Ramp()
    p = np.asarray(params, np.float32)
    a = np.asarray(alim, np.float32)
    ob.lib().orc_ramp_set(C.byref(r), ob.fp(p), ob.fp(a))
    return r


def _vel(r, t):
    v = np.zeros(3, np.float32)
    ob.lib().orc_ramp_velocity(C.byref(r), C.c_float(t), ob.fp(v))
    return v


def test_ramp_hand_derived():
    # SURVEY.md §8c item 1 (RampVelocityControl.cpp:98-158)
    h = KAT["hand"]["ramp"]
    r = _ramp(h["params"], h["alim"])
    assert r.t_acc == pytest.approx(0.3, abs=1e-7)
    assert ob.lib().orc_ramp_max_duration(C.byref(r)) == pytest.approx(1.1, abs=1e-6)
    assert _vel(r, 0.15) == pytest.approx(h["v_at_0.15"], abs=1e-7)
    assert _vel(r, 0.5) == pytest.approx(h["v_at_0.5"], abs=1e-7)
    assert _vel(r, 0.9) == pytest.approx(h["v_at_0.9"], abs=1e-6)
    assert _vel(r, 1.2) == pytest.approx(h["v_at_1.2"], abs=0)


def test_ramp_golden_vectors_bit_exact():
    for case in KAT["ramp"]:
        r = _ramp(case["params"], case["alim"])
        assert r.t_acc == np.float32(case["t_acc"])
        assert [r.a[i] for i in range(3)] == [np.float32(x) for x in case["acc"]]
        assert ob.lib().orc_ramp_max_duration(C.byref(r)) == np.float32(case["T"])
        for t, v in zip(case["ts"], case["vel"]):
            assert list(_vel(r, t)) == [np.float32(x) for x in v], (case["params"], t)


def test_zero_velocity_control_has_zero_accelerations():
    # RampVelocityControl.cpp:151-156
    r = _ramp([0, 0, 0, 0.4, 0], [2, 2, 4])
    assert r.t_acc == 0 and [r.a[i] for i in range(3)] == [0, 0, 0]
    assert ob.lib().orc_ramp_max_duration(C.byref(r)) == np.float32(0.4)


def test_parameter_packing_and_local_twist():
    # (v..., t_plateau, t_rest) RampVelocityControl.cpp:34-61; setFromLocalTwist :68-79
    pose = np.array([0.1, 0.2, 0.5], np.float32)
    tw = np.array([0.3, 0.7, 0.4, 0.25], np.float32)
    out = np.zeros(5, np.float32)
    ob.lib().orc_ramp_from_local_twist(ob.fp(pose), ob.fp(tw), ob.fp(out))
    heading = np.float32(0.5) + np.float32(0.25)
    assert out[0] == pytest.approx(0.3 * math.cos(heading), rel=1e-6)
    assert out[1] == pytest.approx(0.3 * math.sin(heading), rel=1e-6)
    assert out[2] == np.float32(0.7) and out[3] == np.float32(0.4) and out[4] == 0


def test_active_step_count_and_free_motion():
    """#steps = #{k: k*dt (float accumulated) < T}; a robot that touches nothing moves by the left Riemann
    sum of the ramp (SimEnvStatePropagator.cpp:69-83)."""
    sc, _ = scenes.make_scene(1, seed=0, walls=False)
    start = np.array([0.0, 0.0, 0.0, 50.0, 50.0, 0.0], np.float32)
    sc.x_min = sc.y_min = -1e6
    sc.x_max = sc.y_max = 1e6
    for case in KAT["steps"]:
        assert np.allclose(case["alim"], [case["alim"][0]] * 0 + case["alim"])
        for i in range(3):
            sc.accel_limits[i] = case["alim"][i]
        ok, out, c, fl, steps = ob.propagate(sc, start, np.array(case["params"], np.float32))
        rest_steps = round(float(c[4]) / 0.01)
        assert steps - rest_steps == case["n_active"], case
        assert ok
        # positions accumulate x += dt * v in float, exactly like the golden loop
        assert out[0] == np.float32(case["free_disp"][0]) and out[1] == np.float32(case["free_disp"][1])
        th = ob_normalize(float(np.float32(case["free_disp"][2])))
        assert abs(out[2] - th) < 2e-6
        assert out[3] == 50.0 and out[4] == 50.0  # the untouched body does not move


def ob_normalize(v):
    x = C.c_double(v)
    ob.lib().orc_normalize_orientation_d(C.byref(x))
    return x.value


def test_so2_helpers():
    for case in KAT["so2"]:
        assert ob.lib().orc_shortest_direction_so2(case["a"], case["b"]) == np.float32(case["d"])
    for case in KAT["normalize"]:
        assert ob_normalize(case["v"]) == case["n"]
    # edges at +-pi (Math.cpp:8-16): [-pi, pi)
    assert ob_normalize(math.pi) == -math.pi
    assert ob_normalize(-math.pi) == -math.pi
    f = C.c_float(math.pi)
    ob.lib().orc_normalize_orientation_f(C.byref(f))
    assert f.value == pytest.approx(-math.pi, abs=1e-6)


def test_distance_hand_and_golden():
    sc, _ = scenes.make_scene(1, seed=0)
    h = KAT["hand"]["distance_two_bodies"]
    assert ob.distance(sc, h["a"], h["b"]) == pytest.approx(h["all"], abs=1e-12)
    assert ob.distance(sc, h["a"], h["b"], mask=1) == pytest.approx(h["first_only"], abs=1e-12)
    assert ob.distance(sc, h["a"], h["b"], mask=2) == pytest.approx(h["second_only"], abs=1e-12)
    for case in KAT["distance"]:
        sc2 = scenes.default_scene()
        sc2.n_bodies = case["B"]
        d = ob.distance(sc2, case["a"], case["b"], weights=case["w"], mask=case["mask"])
        assert d == case["d"], case["B"]  # same double operation sequence -> identical bits
        # the product's host helper (mps_distance) is the same expression
        out = C.c_double()
        a = np.asarray(case["a"], np.float32)
        b = np.asarray(case["b"], np.float32)
        w = np.asarray(case["w"], np.float32)
        rc = abi.load_library().mps_distance(C.byref(sc2), ob.fp(a), ob.fp(b), ob.fp(w), case["mask"], C.byref(out))
        assert rc == 0 and out.value == case["d"]


def test_goal_distance_golden_and_boundary():
    for case in KAT["goal"]:
        s = np.asarray(case["state"], np.float32)
        ng = len(case["goals"])
        g = (abi.Goal * ng)()
        for i, (body, gx, gy, tol) in enumerate(case["goals"]):
            g[i].body, g[i].x, g[i].y, g[i].tol = body, gx, gy, tol
        assert ob.lib().orc_goal_distance(ob.fp(s), g, ng) == case["d"]
        assert bool(ob.lib().orc_goal_satisfied(ob.fp(s), g, ng)) == case["satisfied"]
    # on the boundary of the tolerance disc: distance exactly 0 -> satisfied (ObjectsRelocationGoal.cpp:113-122)
    s = np.array([0, 0, 0, 0.5, 0.0, 0.0], np.float32)
    g = (abi.Goal * 1)()
    g[0].body, g[0].x, g[0].y, g[0].tol = 1, 0.0, 0.0, 0.5
    assert ob.lib().orc_goal_satisfied(ob.fp(s), g, 1) == 1
    g[0].tol = np.float32(0.4999)
    assert ob.lib().orc_goal_satisfied(ob.fp(s), g, 1) == 0


def test_collision_policy_truth_table():
    # CollisionPolicy::collisionAllowed SimEnvState.cpp:1375-1394 and the PlanningProblem defaults
    # (static collisions allowed, robot-static forbidden: OraclePushPlanner.cpp:90-91)
    sc, _ = scenes.make_scene(3, seed=1)
    B = sc.n_bodies
    L = ob.lib()
    assert L.orc_collision_allowed(C.byref(sc), 0, B + 0) == 0   # robot - static
    assert L.orc_collision_allowed(C.byref(sc), B + 1, 0) == 0   # order does not matter
    assert L.orc_collision_allowed(C.byref(sc), 1, B + 2) == 1   # object - static
    assert L.orc_collision_allowed(C.byref(sc), 0, 1) == 1       # robot - object: unknown pairs are allowed
    assert L.orc_collision_allowed(C.byref(sc), B, B + 1) == 1   # static - static
    sc.pair_forbidden[2] |= 1 << 3
    sc.pair_forbidden[3] |= 1 << 2
    assert L.orc_collision_allowed(C.byref(sc), 2, 3) == 0 and L.orc_collision_allowed(C.byref(sc), 3, 2) == 0
    assert L.orc_collision_allowed(C.byref(sc), 1, 3) == 1


def test_derived_constants():
    # SliceBasedOracleRRT min slice distance (RRT.cpp:780) and HumanOracle::getMaximalPushingDistance (:161-164)
    assert KAT["hand"]["slice_threshold_B11"] == pytest.approx(0.010)
    assert KAT["hand"]["max_pushing_distance"] == pytest.approx(0.6413, abs=1e-4)
