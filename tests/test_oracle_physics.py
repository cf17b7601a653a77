"""CPU tests of the physics spec mps-planar-v3 as implemented by the oracle: mechanics sanity, the
propagate contract of the reference (SimEnvStatePropagator.cpp:42-118) and best-of-K semantics
(NaiveControlSampler.cpp:34-71, RRT.cpp:449-489)."""
import ctypes as C

import numpy as np
import pytest

from manipulation_planning_suite_b200 import abi, scenes

import oracle_binding as ob


def _scene(n=3, **kw):
    sc, _ = scenes.make_scene(n, seed=3, **kw)
    return sc


def test_sincos_accuracy_and_identity():
    xs = np.linspace(-3.2, 3.2, 2001, dtype=np.float32)
    s = C.c_float()
    c = C.c_float()
    for x in xs:
        ob.lib().orc_sincos(C.c_float(x), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(np.float64(x))) < 3e-7
        assert abs(c.value - np.cos(np.float64(x))) < 3e-7


def test_friction_stopping_distance():
    """A sliding object stops after ~ v^2 / (2 mu g) (support friction, no contacts)."""
    sc = _scene(1, walls=False)
    pose = np.array([5.0, 5.0, 0.0, 0.0, 0.0, 0.0], np.float32)
    vel = np.array([0, 0, 0, 0.5, 0.0, 0.0], np.float32)
    u = np.zeros(3, np.float32)
    nm = C.c_int32()
    for _ in range(400):
        ob.lib().orc_step_once(C.byref(sc), ob.fp(pose), ob.fp(vel), ob.fp(u), C.byref(nm))
    assert vel[3] == 0.0 and vel[4] == 0.0
    expect = 0.5 ** 2 / (2 * 0.19 * 9.81)
    assert pose[3] == pytest.approx(expect, rel=0.05)


def test_spinning_object_stops():
    sc = _scene(1, walls=False)
    pose = np.array([5.0, 5.0, 0.0, 0.0, 0.0, 0.0], np.float32)
    vel = np.array([0, 0, 0, 0.0, 0.0, 3.0], np.float32)
    u = np.zeros(3, np.float32)
    nm = C.c_int32()
    for _ in range(400):
        ob.lib().orc_step_once(C.byref(sc), ob.fp(pose), ob.fp(vel), ob.fp(u), C.byref(nm))
    assert vel[5] == 0.0


def test_central_push_moves_object_with_robot():
    sc = _scene(3)
    start = np.array([0.0, 0.0, 0.0, 0.16, 0.0, 0.0, 0.6, 0.6, 0.0, -0.6, -0.6, 0.0], np.float32)
    ctrl = np.array([0.3, 0.0, 0.0, 0.8, 0.0], np.float32)
    ok, out, c, fl, steps = ob.propagate(sc, start, ctrl)
    assert ok and (fl & abi.MPS_F_OK)
    # the robot travelled v * (t_acc + t_plateau) = 0.3 * (0.15 + 0.8); the object stayed in front of it
    assert out[0] == pytest.approx(0.3 * 0.95, abs=0.01)
    gap = out[3] - out[0]
    assert 0.10 < gap < 0.16
    assert abs(out[4]) < 0.02          # central push: no sideways drift
    assert np.array_equal(out[6:], start[6:])  # untouched bodies are bit-identical
    assert c[4] >= 0.0 and c[4] < 1.0  # rest time written back into the control


def test_object_stops_at_wall_and_robot_wall_contact_is_invalid():
    sc = _scene(3)
    # object pushed into the +x wall, robot follows: the robot ends up wedging the object
    start = np.array([0.6, 0.0, 0.0, 0.78, 0.0, 0.0, -0.5, 0.5, 0.0, -0.5, -0.5, 0.0], np.float32)
    ok, out, c, fl, steps = ob.propagate(sc, start, np.array([0.5, 0, 0, 0.9, 0], np.float32))
    assert not ok
    assert fl & (abi.MPS_F_BOUNDS | abi.MPS_F_INFEASIBLE | abi.MPS_F_BAD_CONTACT | abi.MPS_F_JAMMED)
    # a robot parked in the wall is an invalid state (robot-static forbidden)
    bad = start.copy()
    bad[0] = 0.97
    valid, f = ob.is_valid(sc, bad)
    assert not valid and (f & abi.MPS_F_BAD_CONTACT)


def test_bounds_are_checked_before_collisions():
    sc = _scene(3)
    s = np.array([0.0, 0.0, 0.0, 1.5, 0.0, 0.0, -0.5, 0.5, 0.0, -0.5, -0.5, 0.0], np.float32)
    valid, f = ob.is_valid(sc, s)
    assert not valid and f == abi.MPS_F_BOUNDS


def test_overlap_beyond_pen_max_is_infeasible():
    sc = _scene(3)
    s = np.array([0.5, -0.5, 0.0, 0.0, 0.0, 0.0, 0.05, 0.0, 0.0, -0.5, -0.5, 0.0], np.float32)
    valid, f = ob.is_valid(sc, s)
    assert not valid and (f & abi.MPS_F_INFEASIBLE)
    ab, data = ob.contacts(sc, s)
    assert len(ab) == 1 and tuple(ab[0]) == (1, 2)


def test_propagate_is_deterministic_and_alias_safe():
    # the reference's only statement about propagate: deterministic in (state, control) (verifySolution,
    # OraclePushPlanner.cpp:495-519); in/out may alias (:510-512)
    sc = _scene(3)
    start = scenes.cluttered_start(sc, 5)
    ctrl = np.array([0.35, 0.1, 0.4, 0.5, 0.0], np.float32)
    a = ob.propagate(sc, start, ctrl)
    b = ob.propagate(sc, start, ctrl)
    assert np.array_equal(a[1], b[1]) and a[3] == b[3] and np.array_equal(a[2], b[2])
    s = start.copy()
    c = ctrl.copy()
    fl = C.c_uint32()
    st = C.c_int32()
    ok = ob.lib().orc_propagate(C.byref(sc), ob.fp(s), ob.fp(c), ob.fp(s), C.byref(fl), C.byref(st))
    assert bool(ok) == a[0] and np.array_equal(s, a[1])


def test_rest_time_is_reset_and_rewritten():
    # SimEnvStatePropagator.cpp:63-67: the incoming rest time is ignored and overwritten
    sc = _scene(3)
    start = scenes.cluttered_start(sc, 5)
    c1 = np.array([0.35, 0.1, 0.4, 0.5, 0.0], np.float32)
    c2 = c1.copy()
    c2[4] = 3.0
    r1 = ob.propagate(sc, start, c1)
    r2 = ob.propagate(sc, start, c2)
    assert np.array_equal(r1[1], r2[1]) and r1[2][4] == r2[2][4]


def test_best_of_k_ties_first_index_and_all_invalid():
    # NaiveControlSampler.cpp:42-70
    sc = _scene(3)
    start = np.array([0.0, 0.0, 0.0, 0.4, 0.4, 0.0, -0.4, 0.4, 0.0, -0.4, -0.4, 0.0], np.float32)
    c = np.array([0.1, 0.05, 0.2, 0.3, 0.0], np.float32)
    r = ob.extend(sc, start, np.tile(c, (5, 1)), dest=start)
    assert r["k"] == 0 and r["n_ok"] == 1
    bad_start = np.array([0.9, 0.0, 0.0, 0.4, 0.4, 0.0, -0.4, 0.4, 0.0, -0.4, -0.4, 0.0], np.float32)
    r = ob.extend(sc, bad_start, np.tile(np.array([0.6, 0, 0, 1.0, 0], np.float32), (4, 1)), dest=start)
    assert r["k"] == -1 and r["n_ok"] == 0 and np.isinf(r["distance"])
    assert (r["all_flags"] & abi.MPS_F_OK == 0).all()


def test_best_is_argmin_of_masked_distance():
    sc = _scene(3)
    start = scenes.cluttered_start(sc, 9)
    rng = np.random.RandomState(0)
    controls = scenes.sample_controls_towards(rng, sc, start, 24, 1)
    dest = scenes.sample_state(rng, sc)
    for mask in (None, 1 << 1, 0b1110):
        r = ob.extend(sc, start, controls, dest=dest, mask=mask)
        d = np.array([ob.distance(sc, r["all_states"][k, 0], dest, mask=mask) if (r["all_flags"][k, 0] & abi.MPS_F_OK)
                      else np.inf for k in range(24)])
        assert np.array_equal(d, r["all_dist"])
        assert r["k"] == int(np.argmin(d))


def test_chain_aborts_at_first_failure_and_keeps_prefix():
    # HybridActionRRT::forwardPropagateActionSequence RRT.cpp:531-565
    sc = _scene(3)
    start = np.array([0.0, 0.0, 0.0, 0.4, 0.4, 0.0, -0.4, 0.4, 0.0, -0.4, -0.4, 0.0], np.float32)
    chain = np.array([[[0.3, 0.0, 0.0, 0.5, 0.0],    # fine
                       [0.6, 0.0, 0.0, 1.0, 0.0],    # drives the robot out through the wall -> fails
                       [0.0, 0.3, 0.0, 0.5, 0.0]]], np.float32)
    r = ob.extend(sc, start, chain, dest=start, compare_f32=True)
    fl = r["all_flags"][0]
    assert fl[0] & abi.MPS_F_OK and (fl[1] & abi.MPS_F_RAN) and not (fl[1] & abi.MPS_F_OK) and fl[2] == 0
    assert r["n_ok"] == 1 and r["k"] == 0
    # the distance is taken at the last successful state of the prefix
    assert r["all_dist"][0] == ob.distance(sc, r["all_states"][0, 0], start)


def test_multithreaded_extend_equals_sequential():
    sc = _scene(6)
    start = scenes.cluttered_start(sc, 2)
    rng = np.random.RandomState(2)
    controls = scenes.sample_controls_towards(rng, sc, start, 64, 2)
    dest = scenes.sample_state(rng, sc)
    a = ob.extend(sc, start, controls, dest=dest, compare_f32=True)
    b = ob.extend(sc, start, controls, dest=dest, compare_f32=True, threads=4)
    assert a["k"] == b["k"] and a["distance"] == b["distance"]
    assert np.array_equal(a["all_states"], b["all_states"]) and np.array_equal(a["best_states"], b["best_states"])


def test_nn_linear_matches_numpy_bruteforce():
    sc = _scene(4)
    B = sc.n_bodies
    rng = np.random.RandomState(4)
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(500)])
    q = scenes.sample_state(rng, sc)
    d = np.array([ob.distance(sc, s, q) for s in states])
    i, di = ob.nn_linear(sc, states, q)
    assert i == int(np.argmin(d)) and di == d.min()
    dup = np.concatenate([states, states[:10]])
    i, _ = ob.nn_linear(sc, dup, states[3])
    assert i == 3
    i, _ = ob.nn_linear(sc, states, q, slice_ids=np.zeros(500, np.int32), slice_filter=1)
    assert i == -1
    i, _ = ob.nn_linear(sc, states[:0], q)
    assert i == -1


def test_oracle_reproduces_committed_physics_golden():
    """Regression pin of the spec version: the committed vectors (tests/golden/make_physics_golden.py, frozen per
    spec version) must be reproduced bit for bit by the oracle built from source on this machine."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "physics_golden_v3.npz"))
    i = 0
    while f"c{i}_meta" in g:
        n_objects, seed, k, n_ok = [int(x) for x in g[f"c{i}_meta"]]
        sc, _ = scenes.make_scene(n_objects, seed=seed)
        r = ob.extend(sc, g[f"c{i}_start"], g[f"c{i}_controls"], dest=g[f"c{i}_dest"], compare_f32=True)
        assert r["k"] == k and r["n_ok"] == n_ok
        for key in ("all_states", "all_rest", "all_flags", "all_dist", "all_steps", "best_states"):
            assert np.array_equal(r[key], g[f"c{i}_{key}"]), (i, key)
        i += 1
    assert i == 4


# ---- the independent pin of the physics spec -------------------------------------------------------------------
# tests/golden/physics_kat_v3.json comes from tests/golden/spec_np.py, a numpy-float32 restatement of DESIGN.md §3
# that shares no code with the oracle.  The oracle has to reproduce every number of it bit for bit.

def test_fma32_of_the_restatement_is_the_ieee_fused_multiply_add():
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import spec_np as sp

    libm = C.CDLL("libm.so.6")
    libm.fmaf.restype = C.c_float
    libm.fmaf.argtypes = [C.c_float] * 3
    rng = np.random.RandomState(0)
    for i in range(30000):
        a, b = rng.standard_normal(2).astype(np.float32)
        if i % 3 == 0:
            c = np.float32(-(float(a) * float(b)))          # cancellation
        elif i % 3 == 1:
            c = np.float32(rng.standard_normal() * 10.0 ** rng.randint(-8, 8))
        else:
            c = np.float32(rng.standard_normal())           # the product close to half an ulp of c
            b = np.float32(float(np.spacing(c)) * 0.5 / float(a))
        assert sp.fma32(a, b, c) == np.float32(libm.fmaf(a, b, c)), (a, b, c)
    # exact ties of the double sum that the remainder has to break
    one, eps = np.float32(1.0), np.float32(2.0 ** -24)
    for a, b, c in [(one + np.float32(2.0 ** -23), eps, one), (eps, eps, one), (np.float32(3.0), eps, one),
                    (np.float32(-3.0), eps, one), (eps, -eps, one + np.float32(2.0 ** -23))]:
        assert sp.fma32(a, b, c) == np.float32(libm.fmaf(a, b, c)), (a, b, c)


def _kat():
    import physics_kat as pk

    return pk, pk.load()


def test_oracle_reproduces_the_independent_propagate_answers():
    pk, kat = _kat()
    assert kat["spec"] == "mps-planar-v3"
    names = set()
    for case in kat["propagate"]:
        sc = pk.build_scene(case["scene"])
        ok, out, c, fl, steps = ob.propagate(sc, pk.arr(case["start"]), pk.arr(case["control"]))
        assert bool(ok) == case["ok"] and fl == case["flags"] and steps == case["steps"], case["name"]
        assert c[4] == np.float32(case["rest"]), case["name"]
        assert np.array_equal(out, pk.arr(case["out"])), (case["name"], out, case["out"])
        names.add(case["name"])
    assert {"disc_disc_central_push", "box_box_face_two_points", "box_disc_corner", "push_release_slide", "jam_rule",
            "train_box_disc_box", "box_slides_along_wall"} <= names


def test_oracle_reproduces_the_independent_single_step_answers():
    """pose and twist of every body after each of the first steps, through the oracle's single-step entry"""
    pk, kat = _kat()
    n_steps = 0
    for case in kat["propagate"]:
        sc = pk.build_scene(case["scene"])
        pose = pk.arr(case["start"]).copy()
        for i in range(sc.n_bodies):  # set_state wraps theta; the starts are inside [-pi, pi)
            assert -np.pi <= pose[3 * i + 2] < np.pi
        vel = np.zeros_like(pose)
        nm = C.c_int32()
        for t in case["trace"]:
            u = pk.arr(t["u"])
            ob.lib().orc_step_once(C.byref(sc), ob.fp(pose), ob.fp(vel), ob.fp(u), C.byref(nm))
            assert nm.value == t["n_manifolds"], case["name"]
            assert np.array_equal(pose, pk.arr(t["pose"])), (case["name"], n_steps)
            assert np.array_equal(vel, pk.arr(t["twist"])), (case["name"], n_steps)
            n_steps += 1
    assert n_steps > 900


def test_oracle_reproduces_the_independent_validity_and_sincos_answers():
    pk, kat = _kat()
    sc = pk.build_scene(kat["valid"]["scene"])
    for case in kat["valid"]["cases"]:
        valid, f = ob.is_valid(sc, pk.arr(case["state"]))
        assert f == case["flags"] and bool(valid) == (case["flags"] == 0), case["name"]
    s = C.c_float()
    c = C.c_float()
    for case in kat["sincos"]:
        ob.lib().orc_sincos(C.c_float(case["a"]), C.byref(s), C.byref(c))
        assert np.float32(s.value) == np.float32(case["s"]) and np.float32(c.value) == np.float32(case["c"])


def test_restatement_run_live_agrees_with_the_oracle_on_a_fresh_case():
    """the committed file is what spec_np.py computes today, and a case outside the file agrees as well"""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import spec_np as sp

    pk, kat = _kat()
    case = kat["propagate"][2]
    sc = pk.build_scene(case["scene"])
    ok, out, rest, fl, steps = sp.propagate(sc, pk.arr(case["start"]), pk.arr(case["control"]))
    assert np.array_equal(out, pk.arr(case["out"])) and fl == case["flags"] and steps == case["steps"]
    sc, _ = scenes.make_scene(5, seed=77)
    start = scenes.cluttered_start(sc, 77)
    rng = np.random.RandomState(77)
    controls = scenes.sample_controls_towards(rng, sc, start, 3, 1)
    for k in range(3):
        ok, out, c, fl, steps = ob.propagate(sc, start, controls[k, 0])
        ok2, out2, rest2, fl2, steps2 = sp.propagate(sc, start, controls[k, 0])
        assert (bool(ok), fl, steps, c[4]) == (ok2, fl2, steps2, rest2) and np.array_equal(out, out2)


def test_gnat_restatement_is_exact():
    """oracle/mps_gnat.c (ompl::NearestNeighborsGNAT restated) against the linear scan: same distance, and the
    same index whenever the minimum is unique; far fewer distance evaluations than a scan."""
    sc = _scene(6)
    rng = np.random.RandomState(12)
    N = 6000
    states = np.stack([scenes.sample_state(rng, sc) for _ in range(N)])
    for mask in (None, 0b1111110, 1):
        g = ob.Gnat(sc, mask=mask)
        g.add_many(states)
        e0 = g.distance_evaluations()
        for _ in range(40):
            q = scenes.sample_state(rng, sc)
            gi, gd = g.nearest(q)
            wi, wd = ob.nn_linear(sc, states, q, mask=mask)
            assert gd == wd
            if mask is None:
                assert gi == wi
        if mask is None:
            assert (g.distance_evaluations() - e0) / 40 < N
    g = ob.Gnat(sc)
    assert g.nearest(states[0]) == (-1, np.inf)
    dup = np.concatenate([states[:300], states[:300]])
    g.add_many(dup)
    i, d = g.nearest(states[17])
    assert d == 0.0 and i in (17, 317)
