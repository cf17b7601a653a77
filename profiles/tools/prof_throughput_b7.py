import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from manipulation_planning_suite_b200 import api, scenes
n_objects, K = int(os.environ.get("NOBJ", "6")), 32768
sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
start = scenes.cluttered_start(sc, 1000 + n_objects)
rng = np.random.RandomState(n_objects)
controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
ctx = api.Context(sc)
ctx.extend_upload(start, controls, dest=scenes.sample_state(rng, sc))
for _ in range(2):
    ctx.extend_launch()
ctx.synchronize()
c = ctx.counters()
print("steps per launch", c["steps"] / 2, "candidates/step", c["candidates"] / c["steps"], "touching/step", c["touching"] / c["steps"], "passes/step", c["colour_passes"] / c["steps"])
