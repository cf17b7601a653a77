"""cfg5 sweep (device sampling) timing for the library named by MPS_B200_LIB: 4096 instances, 200 iterations"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from manipulation_planning_suite_b200 import api, multi_seed, scenes
sc, start, goals = bench.cfg5_problem(scenes)
ctx = api.Context(sc)
kw = dict(num_control_samples=10, goal_bias=0.2, target_bias=0.4, device_sampling=True)
multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=4096, max_iterations=3, seed=99, **kw)
t0 = time.perf_counter()
res, _ = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=4096, max_iterations=200, seed=5, **kw)
dt = time.perf_counter() - t0
print(os.environ.get("MPS_B200_LIB", "default"), "cfg5 sweep: %.3f s, %.2f M rollouts/s, solved %.3f" % (dt, res.rollouts / dt / 1e6, (res.solved_iteration >= 0).mean()))
