"""Which chain bounds each cfg2 launch of bench.py's batch pool: kernel time, the slowest chain's cycles, steps,
cycles per step and contact statistics.  python profiles/tools/pool_tail.py [n_batches]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from manipulation_planning_suite_b200 import api
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sc, start, batches, goals = bench.make_workload([0], nb)
ctx = api.Context(sc)
ctx.set_kernel_timing(True)
rows = []
for i, (c, n, d) in enumerate(batches):
    ctx.extend(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
    ctx.kernel_times()
    ctx.extend_upload(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
    ctx.extend_launch()
    r = ctx.extend_fetch(dump=True)
    ms = float(ctx.kernel_times()[-1])
    cyc = ctx.chain_cycles()
    steps = r.all_steps.sum(axis=1)
    k = int(np.argmax(cyc))
    rows.append((ms, i, k, int(cyc[k]), int(steps[k]), cyc[k] / max(steps[k], 1), int(steps.max()), float(cyc.mean())))
rows.sort()
for ms, i, k, cy, st, cps, smax, cmean in rows:
    print(f"batch {i:2d}: kernel {ms:.3f} ms | slowest chain {k:4d}: {cy/1e6:.2f} Mcyc = {cy/1.965e6:.3f} ms, {st} steps, {cps:.0f} cyc/step | longest chain {smax} steps | mean chain {cmean/1.965e6:.3f} ms")
