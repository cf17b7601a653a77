"""single-control rollouts (B = 7, contact-rich start): kernel ms per K for the packed (four chains per warp) and the
one-chain-per-warp builds"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes
n_objects = int(os.environ.get("NOBJ", "6"))
sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
start = scenes.cluttered_start(sc, 1000 + n_objects)
rng = np.random.RandomState(n_objects)
controls = scenes.sample_controls_towards(rng, sc, start, 65536, 1)
dest = scenes.sample_state(rng, sc)
out = {}
for K in (4096, 8192, 16384, 32768, 65536):
    ctx = api.Context(sc)
    ctx.extend_upload(start, controls[:K], dest=dest)
    ctx.extend_launch()
    ctx.synchronize()
    out[K] = round(ctx.time_launches(0, 3) / 3, 4)
    ctx.close()
print({k: os.environ.get(k) for k in ("MPS_ROLLOUT_GROUP", "MPS_ROLLOUT_MINB", "NOBJ")}, out)
