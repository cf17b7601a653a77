import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from manipulation_planning_suite_b200 import api, scenes
n_objects=8; K=40960
sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
st = scenes.cluttered_start(sc, 1000 + n_objects)
rng = np.random.RandomState(n_objects)
controls = scenes.sample_controls_towards(rng, sc, st, K, 1)
ctx = api.Context(sc)
ctx.extend_upload(st, controls, dest=scenes.sample_state(rng, sc))
ctx.extend_launch(); ctx.synchronize()
ctx.extend_launch(); ctx.synchronize()
ctx.close()
