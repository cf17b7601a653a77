"""back-to-back kernel time of mps_tree_nearest_k at 1 M x 11 for the library named by MPS_B200_LIB"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes
sc, _ = scenes.make_scene(10, seed=1003)
ctx = api.Context(sc)
tree = api.NearestNeighbors(ctx)
tree.fill_random(1_000_000, seed=1003)
q = scenes.sample_state(np.random.RandomState(1), sc)
out = {}
tree.nearest_upload(q[None, :]); tree.nearest_launch(); ctx.synchronize()
out["nn1"] = round(ctx.time_launches(1, 20) / 20 * 1e3, 2)
for k in (1, 8, 32):
    try:
        tree.nearest_k(q, k)
    except Exception as e:
        pass
    out[f"k{k}"] = round(ctx.time_launches(2, 20) / 20 * 1e3, 2)
print(os.environ.get("MPS_B200_LIB", "default"), out)
