"""Throughput probe: many single-control rollouts per launch (BASELINE cfg 4 / cfg 5 shapes)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

for n_objects, K in ((16, 65536), (8, 40960), (6, 65536), (3, 65536)):
    sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
    start = scenes.cluttered_start(sc, 1000 + n_objects)
    rng = np.random.RandomState(n_objects)
    controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
    dest = scenes.sample_state(rng, sc)
    ctx = api.Context(sc)
    ctx.extend_upload(start, controls, dest=dest)
    ctx.extend_launch()
    ctx.synchronize()
    ctx.counters()
    iters = 3
    ms = ctx.time_launches(0, iters) / iters
    cnt = ctx.counters()
    print(f"B={n_objects + 1} K={K}: {ms:.2f} ms/launch  {K / ms * 1e3:,.0f} rollouts/s  steps/rollout {cnt['steps'] / iters / K:.1f} "
          f"touching/step {cnt['touching'] / max(cnt['steps'], 1):.2f} colour passes/step {cnt['colour_passes'] / max(cnt['steps'], 1):.2f}",
          flush=True)
    ctx.close()
