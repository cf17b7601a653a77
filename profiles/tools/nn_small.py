import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes
sc, _ = scenes.make_scene(6, seed=1003)
ctx = api.Context(sc)
tree = api.NearestNeighbors(ctx)
rng = np.random.RandomState(1)
for N in (64, 1000, 10000, 100000, 1000000):
    tree.clear(); tree.fill_random(N, seed=3)
    q = scenes.sample_state(rng, sc)
    tree.nearest_upload(q[None, :])
    for _ in range(5): tree.nearest_launch()
    ctx.synchronize()
    ms = ctx.time_launches(1, 200) / 200
    t0 = time.perf_counter()
    for _ in range(200): tree.nearest(q)
    e2e = (time.perf_counter() - t0) / 200
    print(f"N={N}: kernel {ms*1e3:.2f} us back-to-back, host e2e {e2e*1e6:.1f} us/query", flush=True)
