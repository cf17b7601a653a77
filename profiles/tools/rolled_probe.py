"""The unpacked 128-register build with the solver sweeps inlined per round against rolled over the rounds
(MPS_ROLLED_MAX_BODIES = 0 / 32), launches that fill the machine, scenes of 9 .. 17 bodies: M rollouts/s."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

out = []
for n_objects, K, L in ((8, 40960, 1), (10, 40960, 1), (12, 40960, 1), (14, 40960, 1), (16, 65536, 1), (6, 16384, 4), (8, 16384, 2)):
    sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
    st = scenes.cluttered_start(sc, 1000 + n_objects)
    rng = np.random.RandomState(n_objects)
    controls = scenes.sample_controls_towards(rng, sc, st, K, L)
    ctx = api.Context(sc)
    ctx.extend_upload(st, controls, dest=scenes.sample_state(rng, sc))
    ctx.extend_launch()
    ctx.synchronize()
    ctx.counters()
    ms = ctx.time_launches(0, 3) / 3
    c = ctx.counters()
    out.append(f"B={n_objects + 1} K={K} L={L}: {K * L / ms / 1e3:.2f} (touching/step {c['touching'] / c['steps']:.2f}, passes/step "
               f"{c['colour_passes'] / c['steps']:.1f}, steps/rollout {c['steps'] / 3 / (K * L):.0f})")
    ctx.close()
print(os.environ.get("MPS_ROLLED_MAX_BODIES", "default"), " | ".join(out))
