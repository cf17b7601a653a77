"""Small profiling workload (run under ncu): 2 cfg2 extends (K=1024 chains), 3 cfg3 NN scans (1M nodes) and two exact
top-k queries (k = 8, 32) over the same tree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

K = int(os.environ.get("PROF_K", "1024"))
sc, start, batches, goals = bench.make_workload([0], 1)
c, n, d = batches[0]
ctx = api.Context(sc)
ctx.extend_upload(start, c[:K], n_ctrl=n[:K], dest=d, compare_f32=True, goals=goals)
for _ in range(2):
    ctx.extend_launch()
r = ctx.extend_fetch()
print("extend winner", r.k, r.distance)
ctx.close()

sc3, _ = scenes.make_scene(bench.NN_CFG["n_objects"], seed=bench.NN_CFG["seed"])
ctx3 = api.Context(sc3)
tree = api.NearestNeighbors(ctx3)
tree.fill_random(bench.NN_CFG["nodes"], seed=bench.NN_CFG["seed"], n_slices=64)
q = scenes.sample_state(np.random.RandomState(11), sc3)
tree.nearest_upload(q[None, :])
for _ in range(3):
    tree.nearest_launch()
print("nn", tree.nearest_fetch())
for k in (8, 32):
    print("top-k", k, tree.nearest_k(q, k)[0][:4])
ctx3.close()
