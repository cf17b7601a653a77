"""Does the grid-stride quantisation (3 vs 4 groups per thread) cost the 1 M-node scan its bandwidth?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes
sc, _ = scenes.make_scene(10, seed=1003)
B = sc.n_bodies
for N in (606208, 909312, 1000000, 1212416, 1515520, 2000000):
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.fill_random(N, seed=1003, n_slices=64)
    q = scenes.sample_state(np.random.RandomState(11), sc)
    tree.nearest_upload(q[None, :])
    for _ in range(5):
        tree.nearest_launch()
    ctx.synchronize()
    ms = ctx.time_launches(1, 100) / 100
    gbs = N * 3 * B * 4 / (ms * 1e-3) / 1e9
    print(f"N={N} groups/thread={N / 4 / (296 * 256):.2f}: {ms * 1e3:7.2f} us {gbs:7.1f} GB/s {gbs / 6554.2:.3f}", flush=True)
    ctx.close()
