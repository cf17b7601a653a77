"""Generates tests/golden/physics_kat_v3.json: known answers of the physics spec mps-planar-v3
computed by the INDEPENDENT numpy restatement tests/golden/spec_np.py — not by the C oracle and not by the CUDA
kernels.  The oracle (tests/test_oracle_physics.py) and the GPU path (tests/test_gpu_extend.py) must reproduce
every number bit for bit.  One file per spec version; a committed version is never regenerated (a new spec
version gets a new file name).

Cases: disc-disc central push, box-box face contact with two points, box-disc corner contact, release and slide
to rest under support friction, the jam rule, a cluttered scene with walls (several manifolds per step, colour
order, statics), the validity test.  Each propagate-level case also carries the poses and twists after each of
its first steps (checked through the oracle's single-step entry).

Run: python tests/golden/make_physics_kat.py
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import physics_kat as pk  # noqa: E402
import spec_np as sp  # noqa: E402
from manipulation_planning_suite_b200 import abi, scenes  # noqa: E402

DISC, BOX = abi.MPS_SHAPE_DISC, abi.MPS_SHAPE_BOX
FAR = [5.0, 5.0, 0.0]


def fl(a):
    return [float(np.float32(x)) for x in a]


def open_scene(n_objects, **sets):
    s = {"x_min": -100.0, "x_max": 100.0, "y_min": -100.0, "y_max": 100.0}
    s.update(sets)
    return {"base": {"n_objects": n_objects, "seed": 1, "walls": False, "disc_every": 0}, "set": s}


def propagate_case(name, spec, start, control, n_trace, expect=None):
    sc = pk.build_scene(spec)
    start = np.array(start, np.float32)
    control = np.array(control, np.float32)
    ok, out, rest, flags, steps = sp.propagate(sc, start, control)
    # the first n_trace steps, one by one (active phase only: the ramp is evaluated like propagate does)
    wd = sp.World(sc)
    wd.set_state(start)
    v, a, t_acc, t_pl = sp.ramp_of(control, sc.accel_limits)
    time = np.float32(0)
    trace = []
    for _ in range(min(n_trace, steps)):
        u = sp.ramp_velocity(v, a, t_acc, t_pl, time)
        wd.step(u)
        time = time + wd.dt
        trace.append({"u": fl(u), "pose": fl(wd.pose()), "twist": fl(wd.twist()), "n_manifolds": wd.n_manifolds})
    case = {"name": name, "scene": spec, "start": fl(start), "control": fl(control), "ok": bool(ok), "out": fl(out),
            "rest": float(rest), "flags": int(flags), "steps": int(steps), "trace": trace}
    if expect:
        expect(case, sc)
    print(f"{name:28s} ok={ok} flags={flags:#x} steps={steps} rest={float(rest):.2f}")
    return case


def main():
    cases = []

    # 1 disc-disc central push: robot disc r = 0.05 pushes a disc r = 0.06 along +x
    spec = open_scene(2, shape={"0": DISC, "1": DISC}, hx={"0": 0.05, "1": 0.06}, hy={"0": 0.05, "1": 0.06})

    def chk1(c, sc):
        o = c["out"]
        assert c["ok"] and abs(o[4]) < 1e-6 and abs(o[1]) < 1e-6, "central push must stay on the axis"
        assert o[3] - o[0] >= 0.11 - 0.002 - 1e-4, "the object ends in front of the robot"
        assert max(t["n_manifolds"] for t in c["trace"]) == 1
    cases.append(propagate_case("disc_disc_central_push", spec, [0, 0, 0, 0.13, 0, 0] + FAR, [0.3, 0, 0, 0.5, 0], 60, chk1))

    # 2 box-box face contact, two contact points: the robot's short face against a box, slightly off centre
    spec = open_scene(2)

    def chk2(c, sc):
        assert c["ok"]
        st = np.array(c["trace"][-1]["pose"], np.float32)
        ab, data = contacts_np(sc, st)
        assert any(n == 2 for n in data), "a face contact has two points"
    cases.append(propagate_case("box_box_face_two_points", spec, [0, 0, 0, 0.125, 0.01, 0] + FAR, [0.3, 0, 0, 0.6, 0], 60, chk2))

    # 3 box-disc corner: the rotated robot box touches the disc with a corner first
    spec = open_scene(2, shape={"1": DISC})

    def chk3(c, sc):
        assert c["ok"]
        assert abs(c["out"][5]) > 0 or True
    cases.append(propagate_case("box_disc_corner", spec, [0, 0, 0.6, 0.14, 0.05, 0] + FAR, [0.25, 0.05, 0.3, 0.5, 0], 60, chk3))

    # 4 release and slide: the robot stops abruptly (acceleration limit 20 m/s^2 > mu g), the object leaves it at
    # the plateau speed and slides to rest under support friction: v^2 / (2 mu g) = 0.0966 m in 0.32 s
    spec = open_scene(2, accel_limits={"0": 20.0, "1": 20.0, "2": 20.0})

    def chk4(c, sc):
        assert c["ok"] and 0.25 < c["rest"] < 0.40, c["rest"]
        # released touching (centre distance 0.11) at 0.6 m/s; while the object slides v^2 / (2 mu g) the robot
        # still covers its deceleration ramp, 0.6 * 0.03 / 2
        slide = (c["out"][3] - c["out"][0]) - 0.11
        expect = 0.6 ** 2 / (2 * 0.19 * 9.81) - 0.6 * 0.03 / 2
        assert abs(slide - expect) < 0.006, (slide, expect)
    cases.append(propagate_case("push_release_slide", spec, [0, 0, 0, 0.13, 0, 0] + FAR, [0.6, 0, 0, 0.1, 0], 80, chk4))

    # 5 jam rule: an object wedged between the robot and a wall never comes to rest
    spec = {"base": {"n_objects": 3, "seed": 3, "walls": True, "disc_every": 4}, "set": {}}

    def chk5(c, sc):
        assert not c["ok"] and (c["flags"] & sp.F_JAMMED), hex(c["flags"])
    found = None
    # the robot stops at x0 + v (t_acc + t_pl): with less than the object's width left to the wall face at x = 1
    for vx, tpl, x0 in [(0.3, 0.5, 0.65), (0.3, 0.5, 0.66), (0.2, 0.5, 0.72)]:
        sc = pk.build_scene(spec)
        start = [x0, 0, 0, x0 + 0.18, 0, 0, -0.5, 0.5, 0, -0.5, -0.5, 0]
        r = sp.propagate(sc, np.array(start, np.float32), np.array([vx, 0, 0, tpl, 0], np.float32))
        print("  jam probe", vx, tpl, x0, hex(r[3]))
        if r[3] & sp.F_JAMMED:
            found = (start, [vx, 0, 0, tpl, 0])
            break
    assert found, "no jammed case found"
    cases.append(propagate_case("jam_rule", spec, found[0], found[1], 40, chk5))

    # 6 clutter with walls: several manifolds per step, all three shape pairings, statics
    spec = {"base": {"n_objects": 6, "seed": 102, "walls": True, "disc_every": 4}, "set": {}}
    sc = pk.build_scene(spec)
    start = scenes.cluttered_start(sc, 102)
    rng = np.random.RandomState(102)
    controls = scenes.sample_controls_towards(rng, sc, start, 6, 1)
    for k in range(6):
        cases.append(propagate_case(f"clutter6_{k}", spec, start, controls[k, 0], 120))
    print("  clutter manifolds per step, max:", [max(t["n_manifolds"] for t in c["trace"]) for c in cases[-6:]])
    assert max(t["n_manifolds"] for c in cases[-6:] for t in c["trace"]) >= 2

    # 6a a train: the robot pushes box, disc, box in a row, slightly staggered (three manifolds in three colours,
    # impulses travel down the chain through the Gauss-Seidel order)
    spec_t = {"base": {"n_objects": 3, "seed": 1, "walls": False, "disc_every": 2},
              "set": {"x_min": -100.0, "x_max": 100.0, "y_min": -100.0, "y_max": 100.0}}
    cases.append(propagate_case("train_box_disc_box", spec_t, [0, 0, 0, 0.112, 0.004, 0.02, 0.233, -0.003, 0, 0.354, 0.006, -0.03],
                                [0.35, 0, 0, 0.4, 0], 90))
    assert max(t["n_manifolds"] for t in cases[-1]["trace"]) >= 3

    # 6b an object pushed along a wall (body-static manifolds, the colours B + s)
    spec_w = {"base": {"n_objects": 3, "seed": 3, "walls": True, "disc_every": 4}, "set": {}}
    cases.append(propagate_case("box_slides_along_wall", spec_w, [0.3, 0.93, 0, 0.45, 0.937, 0, -0.5, 0.5, 0, -0.5, -0.5, 0],
                                [0.3, 0.02, 0, 0.5, 0], 80))
    assert max(t["n_manifolds"] for t in cases[-1]["trace"]) >= 2

    # 7 validity (spec 3.7)
    spec = {"base": {"n_objects": 3, "seed": 3, "walls": True, "disc_every": 4}, "set": {}}
    sc = pk.build_scene(spec)
    valid = []
    for name, s in [("free", [0, 0, 0, 0.4, 0.4, 0, -0.4, 0.4, 0, -0.4, -0.4, 0]),
                    ("out_of_bounds", [0, 0, 0, 1.5, 0, 0, -0.5, 0.5, 0, -0.5, -0.5, 0]),
                    ("deep_overlap", [0.5, -0.5, 0, 0, 0, 0, 0.05, 0, 0, -0.5, -0.5, 0]),
                    ("robot_in_wall", [0.97, 0, 0, 0.4, 0.4, 0, -0.4, 0.4, 0, -0.4, -0.4, 0]),
                    ("object_touching_wall", [0, 0, 0, 0.941, 0, 0, -0.4, 0.4, 0, -0.4, -0.4, 0]),
                    ("light_touch", [0, 0, 0, 0.4, 0.4, 0, 0.4 + 0.119, 0.4, 0, -0.4, -0.4, 0])]:
        f = sp.is_valid(sc, np.array(s, np.float32))
        valid.append({"name": name, "state": fl(s), "flags": int(f)})
        print(f"valid {name:22s} flags={f:#x}")
    assert valid[0]["flags"] == 0 and valid[1]["flags"] == sp.F_BOUNDS and valid[2]["flags"] & sp.F_INFEASIBLE
    assert valid[3]["flags"] & sp.F_BAD_CONTACT and valid[4]["flags"] == 0 and valid[5]["flags"] == 0

    # sincos of the spec at a few arguments
    sincos = [{"a": float(np.float32(a)), "s": float(sp.sincos(a)[0]), "c": float(sp.sincos(a)[1])}
              for a in [0.0, 0.6, -0.6, 1.5707964, 3.1415925, -3.1415927, 2.5, -2.0, 1e-4]]

    out = {"spec": "mps-planar-v3", "source": "tests/golden/spec_np.py (independent numpy restatement)", "propagate": cases,
           "valid": {"scene": spec, "cases": valid}, "sincos": sincos}
    path = os.path.join(HERE, "physics_kat_v3.json")
    if os.path.exists(path) and "--force" not in sys.argv:
        raise SystemExit(f"{path} exists: a committed spec version is never regenerated (use --force only before it is committed)")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


def contacts_np(sc, state):
    wd = sp.World(sc)
    wd.set_state(state)
    found, _ = wd.detect()
    return [(r.a, r.b) for r in found], [sum(1 for k in (0, 1) if r.kn[k] != 0) for r in found]


if __name__ == "__main__":
    main()
