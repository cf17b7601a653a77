"""bench.py's cfg5_sweep leg alone (one GPU): wall time to 90 % solved, rollouts/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from manipulation_planning_suite_b200 import api, scenes  # noqa: E402

torch.cuda.set_device(0)
for _ in range(int(os.environ.get("REPS", "2"))):
    r = bench.bench_cfg5_sweep(api, scenes, torch, None, 0, 0, 1, lambda: torch.cuda.synchronize())
    print(r["wall_s_max_over_ranks"], r["rollouts_per_s"], r["iterations_max_over_ranks"], flush=True)
