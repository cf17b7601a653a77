"""BASELINE config 1: the full planners on a small scene (robot + 3 objects, K = 10), 32 seeds each.
Reports per algorithm: success rate, median time to solution, iterations/s and propagated controls/s
(OraclePushPlanner::solve through the C++ host mirror, every propagation on the device).

usage: python profiles/tools/cfg1_planners.py [seeds] [time_out_s]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import host_api, scenes  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
time_out = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
sc, start = scenes.config_scene(1)
goals = [(1, float(start[3]) + 0.3, float(start[4]) + 0.2, 0.1)]
out = {"config": "cfg1: B=4 (robot + 3 objects), K=10, goal: object 1 moved 0.36 m, tolerance 0.1 m",
       "seeds": n_seeds, "time_out_s": time_out, "algorithms": {}}
for algorithm in ("Naive", "HybridActionRRT", "OracleRRT", "SliceOracleRRT"):
    solved, t_sol, iters, props, wall, path_len, cost0, cost1 = [], [], 0, 0, 0.0, [], [], []
    for seed in range(1, n_seeds + 1):
        t0 = time.perf_counter()
        r = host_api.plan(sc, start, goals, algorithm=algorithm, seed=seed, time_out=time_out, num_control_samples=10,
                          goal_bias=0.1, target_bias=0.1, robot_bias=0.1, shortcut_type="LocalShortcut",
                          max_shortcut_time=5.0)
        dt = time.perf_counter() - t0
        s = r.stats
        solved.append(bool(r.solved))
        iters += s["num_iterations"]
        props += s["num_state_propagations"]
        wall += s["runtime_wall"]
        if r.solved:
            t_sol.append(s["runtime_wall"])
            path_len.append(s["path_len"])
            cost0.append(s["cost_before_shortcut"])
            cost1.append(s["cost_after_shortcut"])
            assert s["reproducible"] == 1 and s["reproducible_after_shortcut"] == 1
    out["algorithms"][algorithm] = {
        "success_rate": round(float(np.mean(solved)), 3),
        "median_time_to_solution_s": None if not t_sol else round(float(np.median(t_sol)), 4),
        "iterations_per_s": round(iters / max(wall, 1e-9), 1),
        "propagations_per_s": round(props / max(wall, 1e-9), 1),
        "median_path_motions_after_shortcut": None if not path_len else float(np.median(path_len)),
        "median_cost_before_after_shortcut_s": None if not cost0 else [round(float(np.median(cost0)), 3),
                                                                      round(float(np.median(cost1)), 3)],
    }
    print(algorithm, out["algorithms"][algorithm], file=sys.stderr, flush=True)
print(json.dumps(out))
