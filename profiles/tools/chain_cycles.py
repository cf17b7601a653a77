"""Which chains are the tail of the cfg2 extend, and what does a physics step cost them?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from manipulation_planning_suite_b200 import api
sc, start, batches, goals = bench.make_workload([0], 1)
c, n, d = batches[0]
ctx = api.Context(sc)
ctx.extend_upload(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
for _ in range(3):
    ctx.extend_launch()
r = ctx.extend_fetch(dump=True)
cyc = ctx.chain_cycles()
steps = r.all_steps.sum(axis=1)
order = np.argsort(-cyc)[:12]
print("kernel-critical chains (cycles, steps, cycles/step):")
for k in order:
    print(f"  chain {k}: {cyc[k]} cycles = {cyc[k]/1.965e3:.0f} us, {steps[k]} steps, {cyc[k]/max(steps[k],1):.0f} cycles/step")
cps = cyc / np.maximum(steps, 1)
print("cycles/step percentiles 10/50/90/99:", np.percentile(cps, [10, 50, 90, 99]).round(0))
print("total cycles", cyc.sum(), "mean chain us", cyc.mean() / 1.965e3)
