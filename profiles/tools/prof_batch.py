"""Profiling workload: two batched NN launches (Q = 32, 1 M nodes x 11 bodies) - the tiled scan"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

sc, _ = scenes.make_scene(bench.NN_CFG["n_objects"], seed=bench.NN_CFG["seed"])
ctx = api.Context(sc)
tree = api.NearestNeighbors(ctx)
tree.fill_random(bench.NN_CFG["nodes"], seed=bench.NN_CFG["seed"], n_slices=64)
rng = np.random.RandomState(11)
qs = np.stack([scenes.sample_state(rng, sc) for _ in range(32)])
tree.nearest_upload(qs)
for _ in range(2):
    tree.nearest_launch()
print(tree.nearest_fetch()[0][:4])
ctx.close()
