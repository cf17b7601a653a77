import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes
n_objects, K = 16, int(os.environ.get("PROF_K", "65536"))
sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
start = scenes.cluttered_start(sc, 1000 + n_objects)
rng = np.random.RandomState(n_objects)
controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
ctx = api.Context(sc)
ctx.extend_upload(start, controls, dest=scenes.sample_state(rng, sc))
for _ in range(2):
    ctx.extend_launch()
ctx.synchronize()
print("done")
