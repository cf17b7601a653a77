import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from manipulation_planning_suite_b200 import host_api, scenes
sc, start = scenes.make_scene(3, seed=1001)
goals = [(1, float(start[3]) + 0.25, float(start[4]) + 0.15, 0.1)]
r = host_api.plan(sc, start, goals, algorithm=sys.argv[1] if len(sys.argv)>1 else "Naive", seed=1, max_iterations=50, time_out=60.0)
print(r.solved, r.stats)
