"""Two chains per warp with the solver sweeps rolled over the rounds (rollout_kernel<4,4,16,true>) against one chain per
warp (the launcher's default), single-control rollouts, scenes of 9 .. 12 bodies, K = 2048 .. 65 536: M rollouts/s.
Run with MPS_ROLLOUT_GROUP=16 MPS_ROLLED_G16=1 for the packed build, without for the default."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

out = []
for n_objects in (8, 9, 10, 11):
    sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
    st = scenes.cluttered_start(sc, 1000 + n_objects)
    row = []
    for K in (2048, 4096, 8192, 16384, 40960, 65536):
        rng = np.random.RandomState(n_objects)
        controls = scenes.sample_controls_towards(rng, sc, st, K, 1)
        ctx = api.Context(sc)
        ctx.extend_upload(st, controls, dest=scenes.sample_state(rng, sc))
        ctx.extend_launch()
        ctx.synchronize()
        ms = ctx.time_launches(0, 3) / 3
        row.append(f"{K}: {K / ms / 1e3:.2f}")
        ctx.close()
    out.append(f"B={n_objects + 1} " + " ".join(row))
print(os.environ.get("MPS_ROLLOUT_GROUP", "default"), " | ".join(out))
