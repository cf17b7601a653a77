"""Latency of the single calls a planner iteration is made of (config 1's scene: robot + 3 objects), through ctypes"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

sc, start = scenes.config_scene(1)
rng = np.random.RandomState(1)
ctx = api.Context(sc)
tree = api.NearestNeighbors(ctx)
tree.add_batch(np.stack([scenes.sample_state(rng, sc) for _ in range(2000)]))
states = np.stack([scenes.sample_state(rng, sc) for _ in range(64)])
controls = scenes.sample_controls_towards(rng, sc, start, 10, 1)


def t(f, n=300):
    for _ in range(20):
        f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e6


print(f"is_valid(1 state)        {t(lambda: ctx.is_valid(states[:1])):7.1f} us")
print(f"is_valid(8 states)       {t(lambda: ctx.is_valid(states[:8])):7.1f} us")
print(f"nearest (2000 nodes)     {t(lambda: tree.nearest(states[3])):7.1f} us")
print(f"extend K=10 L=1          {t(lambda: ctx.extend(start, controls, dest=states[5], mask=2)):7.1f} us")
print(f"tree.add (1 node)        {t(lambda: tree.add(states[7], parent=0)):7.1f} us")
print(f"tree.size                {t(lambda: tree.size()):7.1f} us")
print(f"propagate (K=L=1)        {t(lambda: ctx.propagate(start, controls[0, 0])):7.1f} us")
