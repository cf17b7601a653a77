"""BASELINE config 5 probe: n independent NaiveRRT instances (8 objects, K = 10) on one device."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, multi_seed, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
sc, start = scenes.make_scene(8, seed=1005)
goals = [(1, float(start[3]) + 0.25, float(start[4]) + 0.1, 0.1)]
ctx = api.Context(sc)
res, forest = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, num_control_samples=10, max_iterations=iters,
                                         seed=5, goal_bias=0.2, target_bias=0.4,
                                         device_sampling=os.environ.get("SWEEP_DEVICE", "0") == "1")
solved = (res.solved_iteration >= 0)
print(json.dumps({"instances": n, "iterations": res.iterations, "wall_s": round(res.wall_s, 3), "rollouts": res.rollouts,
                  "rollouts_per_s": round(res.rollouts / res.wall_s, 1), "solved_fraction": round(float(solved.mean()), 4),
                  "median_time_to_solution_s": None if not solved.any() else round(float(np.nanmedian(res.solved_time)), 3),
                  "seconds_in": getattr(res, "timing", None)}))
