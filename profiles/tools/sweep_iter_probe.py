"""per-iteration cost of the device-resident NaiveRRT sweep at a given number of instances (cfg 5 problem)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from manipulation_planning_suite_b200 import api, multi_seed, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
sc, start, goals = bench.cfg5_problem(scenes)
ctx = api.Context(sc)
kw = dict(num_control_samples=10, goal_bias=0.2, target_bias=0.4, device_sampling=True)
multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, max_iterations=3, seed=99, **kw)
t0 = time.perf_counter()
res, _ = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, max_iterations=iters, seed=5, **kw)
dt = time.perf_counter() - t0
print(f"{n} instances: {res.iterations} iterations in {dt:.3f} s = {dt / res.iterations * 1e3:.3f} ms per iteration, device_step {res.timing}")
