"""CPU tests of the host-side C++ mirror (libmps_host.so): closed-form oracle pieces that need no device
(SURVEY.md §8c items 8-9)."""
import numpy as np
import pytest

from manipulation_planning_suite_b200 import host_api, scenes


def test_max_pushing_distance():
    # HumanOracle::getMaximalPushingDistance (oracle/HumanOracle.cpp:161-164): sqrt(2*0.1 + 2*0.1) + 0.01*sqrt(0.78)
    assert host_api.load_host_library().mps_host_max_pushing_distance() == pytest.approx(0.6413, abs=1e-4)


@pytest.mark.parametrize("target", [(1.0, 0.5, 0.3), (0.05, 0.0, 0.0), (-0.7, 0.9, -2.0), (0.0, 0.0, 1.0)])
def test_ramp_computer_steer_reaches_the_target(target):
    """RampComputer::steer (oracle/RampComputer.cpp:40-92): the ramps' areas add up to the straight-line
    displacement, every ramp respects the velocity limits and the plateau limit."""
    vlim, alim = np.array([0.6, 0.6, 1.5]), np.array([2.0, 2.0, 4.0])
    c = host_api.ramp_steer([0, 0, 0], target)
    assert len(c) >= 1
    tot = np.zeros(3)
    for p in c:
        v = p[:3].astype(np.float64)
        assert (np.abs(v) <= vlim + 1e-6).all()
        assert 0.0 <= p[3] <= 1.0 + 1e-6 and p[4] == 0.0
        t_acc = np.max(np.abs(v) / alim)
        tot += v * (t_acc + p[3])
    assert tot == pytest.approx(np.array(target), abs=2e-5)
    # all ramps point along the same direction
    d = np.array(target) / np.linalg.norm(target)
    for p in c:
        assert np.cross(p[:3] / np.linalg.norm(p[:3]), d) == pytest.approx(np.zeros(3), abs=1e-5)


def test_ramp_computer_zero_distance_gives_no_control():
    assert len(host_api.ramp_steer([0.3, 0.2, 0.1], [0.3, 0.2, 0.1])) == 0


def test_planner_without_device_is_a_hard_error():
    import ctypes as C

    from manipulation_planning_suite_b200 import abi

    n = C.c_int(0)
    abi.load_library().mps_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    sc, start = scenes.make_scene(3, seed=1001)
    with pytest.raises(RuntimeError) as e:
        host_api.plan(sc, start, [(1, 0.5, 0.5, 0.1)], algorithm="Naive", max_iterations=5)
    assert "no CPU fallback" in str(e.value)


def test_solution_file_roundtrip_is_lossless(tmp_path):
    """saveSolution / loadSolution format (OraclePushPlanner.cpp:260-385)."""
    from manipulation_planning_suite_b200 import solution_io

    rng = np.random.RandomState(3)
    states = rng.uniform(-1, 1, size=(5, 12)).astype(np.float32)
    controls = rng.uniform(-1, 1, size=(5, 5)).astype(np.float32)
    stats = dict(num_iterations=17, num_state_propagations=170, num_samples=17, num_nearest_neighbor_queries=17,
                 runtime=0.25, success=1, reproducible=1)
    p = tmp_path / "sol.txt"
    solution_io.save_solution(str(p), states, controls, stats)
    lines = p.read_text().split("\n")
    assert lines[0] == "4,0,3,3,3,3" and lines[1] == "5" and lines[3] == "5"
    sol = solution_io.load_solution(str(p), n_objects=4)
    assert np.array_equal(sol.states, states) and np.array_equal(sol.controls, controls)
    assert sol.stats["num_state_propagations"] == 170 and sol.stats["success"] == 1
    with pytest.raises(ValueError):
        solution_io.load_solution(str(p), n_objects=5)


def test_data_triplet_csv_round_trip(tmp_path):
    """OracleDataDumper format (util/Serialize.cpp:73-119): state, result, control[, annotation] per line."""
    from manipulation_planning_suite_b200 import data_generator as dg
    rng = np.random.RandomState(3)
    B = 3
    starts = rng.uniform(-1, 1, size=(7, 3 * B)).astype(np.float32)
    results = rng.uniform(-1, 1, size=(7, 3 * B)).astype(np.float32)
    controls = rng.uniform(-1, 1, size=(7, 5)).astype(np.float32)
    f = tmp_path / "data.csv"
    dg.write_triplets(str(f), starts, results, controls, annotation="box, 0.35")
    lines = f.read_text().splitlines()
    assert len(lines) == 7 and all(len(ln.split(", ")) == 2 * 3 * B + 5 + 2 for ln in lines)
    s, r, c, notes = dg.read_triplets(str(f), B)
    assert np.array_equal(s, starts) and np.array_equal(r, results) and np.array_equal(c, controls)
    assert notes == ["box, 0.35"] * 7
    dg.write_triplets(str(f), starts[:2], results[:2], controls[:2])   # no annotation: no trailing field
    assert len(f.read_text().splitlines()[0].split(", ")) == 2 * 3 * B + 5


def test_host_library_exports_and_struct_layouts():
    """libmps_host.so exports every entry point of host/mps_host_c.h and the ctypes structures match the C ones
    (checked against a C program compiled with gcc)."""
    import ctypes as C
    import os
    import re
    import subprocess
    import tempfile

    L = host_api.load_host_library()
    hdr = os.path.join(os.path.dirname(host_api.HOST_LIB_PATH), "mps_host_c.h")
    names = re.findall(r"\b(mps_host_\w+)\s*\(", open(hdr).read())
    assert {"mps_host_plan", "mps_host_shortcut", "mps_host_default_params", "mps_host_ramp_steer",
            "mps_host_max_pushing_distance"} <= set(names)
    for n in names:
        assert hasattr(L, n), f"{n} not exported"
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu %%zu\\n", sizeof(mps_host_params), sizeof(mps_host_result), offsetof(mps_host_params, shortcut_type), offsetof(mps_host_result, shortcut_accepted));return 0;}\n' % hdr
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-o", exe, c], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(host_api.HostParams), C.sizeof(host_api.HostResult),
                   host_api.HostParams.shortcut_type.offset, host_api.HostResult.shortcut_accepted.offset]
    p = host_api.default_params()
    assert p.shortcut_type == host_api.SHORTCUT_TYPES["LocalShortcut"] and p.max_shortcut_time == 5.0  # reference defaults


def test_project_slice_ball_inside_on_and_outside():
    """SliceBasedOracleRRT::projectSliceBall (host/mps_host.cpp; do_slice_ball_projection of README.md:33-34): a sample
    inside or on the ball of the arrangement metric is left alone; outside, every object moves towards the centre's
    arrangement along the straight line / shortest arc so that the arrangement distance becomes the radius; the robot's
    sub-state is never touched"""
    import oracle_binding as ob

    sc, _ = scenes.make_scene(3, seed=2)
    B = sc.n_bodies
    arr_mask = ((1 << B) - 1) & ~(1 << sc.robot_id)
    rng = np.random.RandomState(0)
    center = scenes.sample_state(rng, sc)
    for trial in range(20):
        sample = scenes.sample_state(rng, sc)
        d = ob.distance(sc, sample, center, mask=arr_mask)
        on = np.float32(d)   # the radius travels as a float: the smallest one that is not below d
        if float(on) < d:
            on = np.nextafter(on, np.float32(np.inf))
        for radius in (d * 1.5, float(on), d * 0.4, d * 0.05):
            inside, out = host_api.project_slice_ball(sc, center, sample, radius)
            assert np.array_equal(out[:3], sample[:3])                      # robot untouched
            if radius >= d:
                assert inside and np.array_equal(out, sample)                # inside / on the ball: unchanged
            else:
                assert not inside
                assert ob.distance(sc, out, center, mask=arr_mask) == pytest.approx(radius, rel=2e-5)
                f = radius / d
                for i in range(1, B):
                    # positions on the segment towards the centre, angles on the shortest arc, by the factor r / d
                    assert out[3 * i:3 * i + 2] == pytest.approx(center[3 * i:3 * i + 2] + f * (sample[3 * i:3 * i + 2] - center[3 * i:3 * i + 2]), abs=1e-6)
                    arc = (sample[3 * i + 2] - center[3 * i + 2] + np.pi) % (2 * np.pi) - np.pi
                    got = (out[3 * i + 2] - center[3 * i + 2] + np.pi) % (2 * np.pi) - np.pi
                    assert got == pytest.approx(f * arc, abs=2e-6)
                    assert -np.pi <= out[3 * i + 2] < np.pi
    # the centre itself: distance 0, inside
    inside, out = host_api.project_slice_ball(sc, center, center, 0.1)
    assert inside and np.array_equal(out, center)
    # per-object weights enter the metric
    w = np.array([1.0, 2.0, 0.5, 1.5], np.float32)
    sample = scenes.sample_state(rng, sc)
    d = ob.distance(sc, sample, center, weights=w, mask=arr_mask)
    inside, out = host_api.project_slice_ball(sc, center, sample, d * 0.5, weights=w)
    assert not inside and ob.distance(sc, out, center, weights=w, mask=arr_mask) == pytest.approx(0.5 * d, rel=2e-5)
