"""Times the rollout kernel of one build of the library (MPS_B200_LIB=...) on the three shapes that matter:
cfg2 extends (K = 1024 chains of <= 4 controls, B = 7: latency regime), config 4's shape (B = 17, K = 65536
single controls) and the same with B = 7 (issue regime).  One JSON line.  Used by variants.sh."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 12
sc, start, batches, goals = bench.make_workload([0], n_batches)
ctx = api.Context(sc)
for i in range(2):
    c, n, d = batches[i]
    ctx.extend(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
ctx.set_kernel_timing(True)
ctx.counters()
rollouts = 0
for i in range(n_batches):
    c, n, d = batches[i]
    r = ctx.extend(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals, dump=True)
    rollouts += int(((r.all_flags & 1) != 0).sum())
kt = ctx.kernel_times()
out = {"lib": os.environ.get("MPS_B200_LIB", "default"), "cfg2_kernel_ms": round(float(kt.mean()), 4),
       "cfg2_Mrollouts_s": round(float(rollouts / kt.sum() / 1e3), 3)}
ctx.close()
for name, n_objects, K in (("cfg4_B17", 16, 65536), ("B7", 6, 65536), ("cfg5_B9", 8, 40960)):
    s2, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
    st = scenes.cluttered_start(s2, 1000 + n_objects)
    rng = np.random.RandomState(n_objects)
    controls = scenes.sample_controls_towards(rng, s2, st, K, 1)
    ctx = api.Context(s2)
    ctx.extend_upload(st, controls, dest=scenes.sample_state(rng, s2))
    ctx.extend_launch()
    ctx.synchronize()
    ms = ctx.time_launches(0, 3) / 3
    out[name + "_Mrollouts_s"] = round(float(K / ms / 1e3), 3)
    ctx.close()
print(json.dumps(out))
