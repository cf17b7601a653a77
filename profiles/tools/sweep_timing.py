import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from manipulation_planning_suite_b200 import api, scenes, multi_seed
sc, start, goals = bench.cfg5_problem(scenes)
ctx = api.Context(sc)
kw = dict(num_control_samples=10, goal_bias=0.2, target_bias=0.4, device_sampling=True)
multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=4096, max_iterations=3, seed=99, **kw)
ctx.set_kernel_timing(True)
res, _ = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=4096, max_iterations=1500, seed=5, stop_fraction=0.9, **kw)
kt = ctx.kernel_times()
print("wall", res.wall_s, "in naive_step calls", res.timing, "rollout kernel sum ms", float(kt.sum()), "launches", len(kt), "iterations", res.iterations)
