"""Generates tests/golden/kat_vectors.json — known-answer vectors for the parts of the path whose
formulas are fully in the reference tree.  The reference itself cannot be built or
imported here (sim_env / OMPL / Eigen absent), so the vectors come from an independent numpy
restatement of those in-tree formulas plus hand-derived constants, NOT from the C oracle:

  ramp profile      src/mps/planner/ompl/control/RampVelocityControl.cpp:98-158
  step count        src/mps/planner/ompl/control/SimEnvStatePropagator.cpp:69-83 (float time accumulation)
  SO(2) helpers     src/mps/planner/util/Math.cpp:8-44
  distance          src/mps/planner/ompl/state/SimEnvWorldStateDistanceMeasure.cpp:34-51 (+ OMPL sub-space semantics)
  goal distance     src/mps/planner/ompl/state/goal/ObjectsRelocationGoal.cpp:108-123

Run: python tests/golden/make_kat.py   (deterministic; rewrites the json next to it)
"""
import json
import math
import os

import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from spec_np import fma32  # noqa: E402  exact binary32 fused multiply-add (checked against libm in the tests)

f32 = np.float32


def ramp(params, alim):
    v = [f32(params[i]) for i in range(3)]
    t_pl = f32(params[3])
    t_acc = f32(abs(v[0] / f32(alim[0])))
    for i in range(3):
        t = f32(abs(v[i] / f32(alim[i])))
        t_acc = t if t_acc < t else t_acc
    if t_acc != 0:
        inv = f32(1.0) / t_acc
        a = [f32(inv * v[i]) for i in range(3)]
    else:
        a = [f32(0)] * 3
    return v, a, t_acc, t_pl


def ramp_velocity(v, a, t_acc, t_pl, t):
    t = f32(t)
    if t < t_acc:
        return [f32(t * a[i]) for i in range(3)]
    if t < f32(t_acc + t_pl):
        return list(v)
    if float(t) < 2.0 * float(t_acc) + float(t_pl):
        s = f32(f32(t - t_acc) - t_pl)
        return [f32(v[i] - f32(s * a[i])) for i in range(3)]
    return [f32(0)] * 3


def max_duration(t_acc, t_pl, t_rest):
    return f32(f32(f32(2.0) * t_acc + t_pl) + f32(t_rest))


def distance(a, b, w, mask, pw=1.0, ow=0.08):
    d = 0.0
    B = len(a) // 3
    for i in range(B):
        if (mask >> i) & 1:
            dx = float(a[3 * i]) - float(b[3 * i])
            dy = float(a[3 * i + 1]) - float(b[3 * i + 1])
            dpos = math.sqrt(dx * dx + dy * dy)
            dd = abs(float(a[3 * i + 2]) - float(b[3 * i + 2]))
            arc = 2.0 * math.pi - dd if dd > math.pi else dd
            d += float(w[i]) * (pw * dpos + ow * arc)
    return d


def goal_distance(state, goals):
    d = 0.0
    for body, gx, gy, tol in goals:
        fx = f32(f32(gx) - f32(state[3 * body]))
        fy = f32(f32(gy) - f32(state[3 * body + 1]))
        dc = math.sqrt(float(fx) * float(fx) + float(fy) * float(fy))
        d += max(dc - float(f32(tol)), 0.0)
    return d


def normalize_d(v):
    v = math.fmod(v, 2.0 * math.pi)
    if v < -math.pi:
        v += 2.0 * math.pi
    elif v >= math.pi:
        v -= 2.0 * math.pi
    return v


def shortest_so2(a, b):
    PI = f32(math.pi)
    v = f32(f32(b) - f32(a))
    if abs(v) > PI:
        v = f32(v - f32(2.0) * PI) if v > 0 else f32(v + f32(2.0) * PI)
    return v


def main():
    rng = np.random.RandomState(20260801)
    out = {"ramp": [], "steps": [], "so2": [], "normalize": [], "distance": [], "goal": []}
    # hand-derived: v = (0.3, 0, 0), a_lim = (1, 1, 1), t_pl = 0.5
    out["hand"] = {
        "ramp": {"params": [0.3, 0.0, 0.0, 0.5, 0.0], "alim": [1.0, 1.0, 1.0], "t_acc": 0.3, "T": 1.1,
                 "v_at_0.15": [0.15, 0.0, 0.0], "v_at_0.9": [0.2, 0.0, 0.0], "v_at_0.5": [0.3, 0.0, 0.0],
                 "v_at_1.2": [0.0, 0.0, 0.0]},
        "max_pushing_distance": math.sqrt(2 * 0.1 + 2 * 0.1) + 0.01 * math.sqrt(0.78),  # HumanOracle.cpp:161-164
        "slice_threshold_B11": 10 * 0.001,  # RRT.cpp:780
        "distance_two_bodies": {"a": [0, 0, 0, 1, 1, 0.5], "b": [3, 4, 0, 1, 1, -0.5],
                                "all": 5.0 + 0.08 * 1.0, "first_only": 5.0, "second_only": 0.08 * 1.0},
    }
    for _ in range(40):
        alim = rng.uniform(0.5, 4.0, size=3)
        params = np.concatenate([rng.uniform(-0.6, 0.6, size=2), rng.uniform(-1.5, 1.5, size=1),
                                 rng.uniform(0.1, 1.0, size=1), [0.0]]).astype(np.float32)
        if _ % 10 == 0:
            params[:3] = 0
        v, a, t_acc, t_pl = ramp(params, alim.astype(np.float32))
        T = max_duration(t_acc, t_pl, 0.0)
        ts = [0.0, float(t_acc) * 0.5, float(t_acc), float(t_acc + t_pl) - 1e-4, float(t_acc + t_pl) + 1e-4, float(T) - 1e-4, float(T) + 0.1]
        out["ramp"].append({"params": [float(x) for x in params], "alim": [float(f32(x)) for x in alim],
                            "t_acc": float(t_acc), "acc": [float(x) for x in a], "T": float(T), "ts": [float(f32(t)) for t in ts],
                            "vel": [[float(x) for x in ramp_velocity(v, a, t_acc, t_pl, t)] for t in ts]})
        # number of physics steps of the active phase: float accumulation of dt = 0.01f (SimEnvStatePropagator.cpp:75-83)
        n = 0
        time = f32(0.0)
        disp = np.zeros(3, np.float32)
        while time < T:
            u = ramp_velocity(v, a, t_acc, t_pl, time)
            for i in range(3):
                disp[i] = fma32(f32(0.01), u[i], disp[i])  # spec v3 §3.2 step 5: x := fma(dt, v, x)
            time = f32(time + f32(0.01))
            n += 1
        out["steps"].append({"params": [float(x) for x in params], "alim": [float(f32(x)) for x in alim], "n_active": n,
                             "free_disp": [float(x) for x in disp]})
    for _ in range(30):
        a, b = rng.uniform(-math.pi, math.pi, size=2)
        out["so2"].append({"a": float(f32(a)), "b": float(f32(b)), "d": float(shortest_so2(a, b))})
    for v in [0.0, 3.0, -3.0, math.pi, -math.pi, 3.5, -3.5, 7.0, -7.0, 100.0, float(f32(math.pi)), float(f32(-math.pi))]:
        out["normalize"].append({"v": v, "n": normalize_d(v)})
    for _ in range(40):
        B = int(rng.randint(2, 18))
        a = rng.uniform(-1, 1, size=3 * B).astype(np.float32)
        b = rng.uniform(-1, 1, size=3 * B).astype(np.float32)
        a[2::3] = rng.uniform(-math.pi, math.pi, size=B)
        b[2::3] = rng.uniform(-math.pi, math.pi, size=B)
        w = rng.uniform(0.2, 3.0, size=B).astype(np.float32) if _ % 2 else np.ones(B, np.float32)
        mask = int(rng.randint(1, 1 << B))
        out["distance"].append({"B": B, "a": [float(x) for x in a], "b": [float(x) for x in b], "w": [float(x) for x in w],
                                "mask": mask, "d": distance(a, b, w, mask)})
    for _ in range(30):
        B = 5
        s = rng.uniform(-1, 1, size=3 * B).astype(np.float32)
        goals = []
        for g in range(int(rng.randint(1, 4))):
            body = int(rng.randint(0, B))
            tol = float(f32(rng.uniform(0.02, 0.2)))
            if rng.rand() < 0.5:  # inside the tolerance disc
                ang, r = rng.uniform(0, 2 * math.pi), tol * rng.uniform(0, 0.95)
                gx, gy = float(s[3 * body]) + r * math.cos(ang), float(s[3 * body + 1]) + r * math.sin(ang)
            else:
                gx, gy = rng.uniform(-1, 1, size=2)
            goals.append([body, float(f32(gx)), float(f32(gy)), tol])
        d = goal_distance(s, goals)
        out["goal"].append({"state": [float(x) for x in s], "goals": goals, "d": d, "satisfied": bool(d <= np.finfo(np.float64).eps)})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
