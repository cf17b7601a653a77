import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from manipulation_planning_suite_b200 import api, multi_seed, scenes
sc, start = scenes.make_scene(8, seed=1005)
goals = [(1, float(start[3]) + 0.25, float(start[4]) + 0.1, 0.1)]
for dev in (False, True):
    ctx = api.Context(sc)
    ctx.counters()
    t0 = time.perf_counter()
    res, forest = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=4096, num_control_samples=10, max_iterations=60,
                                             seed=5, goal_bias=0.2, target_bias=0.4, device_sampling=dev)
    dt = time.perf_counter() - t0
    c = ctx.counters()
    print("device" if dev else "host", f"wall {dt:.3f}s steps {c['steps']} launches {c['launches']} steps/rollout {c['steps']/res.rollouts:.1f} touching/step {c['touching']/max(c['steps'],1):.3f} sizes mean {res.tree_sizes.mean():.2f} timing {res.timing}")
    # time pieces of the device path
    if dev:
        act = np.arange(4096, dtype=np.int32)
        for _ in range(2):
            forest.naive_step(act, 100, 5, goals, K=10, C_=8)
        t0 = time.perf_counter()
        for i in range(10):
            forest.naive_step(act, 200 + i, 5, goals, K=10, C_=8)
        print("naive_step ms", (time.perf_counter() - t0) / 10 * 1e3)
    ctx.close()
