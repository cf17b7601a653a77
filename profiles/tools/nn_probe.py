"""Times the NN scan (cfg3 shape: B=11, 1M nodes) with a given build of libmps_b200.so."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import abi  # noqa: E402

if len(sys.argv) > 1:
    abi._lib = abi.load_library(sys.argv[1])
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

for n_objects, N in ((10, 1_000_000), (6, 100_000), (10, 4_000_000)):
    sc, _ = scenes.make_scene(n_objects, seed=1003)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.fill_random(N, seed=1003, n_slices=64)
    q = scenes.sample_state(np.random.RandomState(11), sc)
    tree.nearest_upload(q[None, :])
    for _ in range(5):
        tree.nearest_launch()
    ctx.synchronize()
    ms = ctx.time_launches(1, 100) / 100
    B = sc.n_bodies
    gbs = N * 3 * B * 4 / (ms * 1e-3) / 1e9
    print(f"{os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else 'default':28s} B={B} N={N}: {ms * 1e3:7.2f} us  {gbs:7.1f} GB/s  {gbs / 6554.2:.3f}", flush=True)
    ctx.close()
