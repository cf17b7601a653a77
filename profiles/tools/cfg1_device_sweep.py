"""BASELINE config 1 (robot + 3 objects, NaiveRRT, K = 10) through the device-resident planner iteration
(mps_forest_naive_step): n independent seeds run side by side in one forest, each iteration = one launch sequence with no
host work but the solved flags.  Compare with profiles/r02_cfg1_planners.json (one planner at a time through the host
mirror: 3.6 k iterations/s, median 0.74 s to a solution).

usage: python profiles/tools/cfg1_device_sweep.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, multi_seed, scenes  # noqa: E402

sc, start = scenes.config_scene(1)
goals = [(1, float(start[3]) + 0.3, float(start[4]) + 0.2, 0.1)]
kw = dict(num_control_samples=10, goal_bias=0.1, target_bias=0.1, robot_bias=0.1, device_sampling=True)
out = {"config": "cfg1: B=4 (robot + 3 objects), NaiveRRT, K=10, goal: object 1 moved 0.36 m, tolerance 0.1 m; "
                 "n seeds side by side in one forest, whole iterations on the device", "runs": []}
ctx = api.Context(sc)
multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=4, max_iterations=5, seed=3, **kw)  # warm-up
for n in (1, 32, 256, 4096):
    res, _ = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, max_iterations=20000 if n <= 32 else 6000, seed=11, **kw)
    t = np.sort(res.solved_time[np.isfinite(res.solved_time)])
    out["runs"].append({"seeds": n, "solved": int(len(t)), "iterations": int(res.iterations), "wall_s": round(res.wall_s, 4),
                        "iterations_per_s": round(res.iterations / res.wall_s, 1),
                        "rollouts_per_s": round(res.rollouts / res.wall_s, 1),
                        "time_to_solution_s": None if not len(t) else {"first": round(float(t[0]), 4), "median": round(float(np.median(t)), 4),
                                                                       "p90": round(float(t[int(0.9 * (len(t) - 1))]), 4), "last": round(float(t[-1]), 4)}})
    print(out["runs"][-1], file=sys.stderr, flush=True)
ctx.close()
print(json.dumps(out))
