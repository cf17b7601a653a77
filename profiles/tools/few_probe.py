"""cfg2-shaped extends with K chains (few-chains regime): kernel ms per K for the grid rule of mps_launch_rollout"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from manipulation_planning_suite_b200 import api
sc, start, batches, goals = bench.make_workload(list(range(int(os.environ.get('SLICES', '2')))), 4)
ctx = api.Context(sc)
out = {}
for K in [int(x) for x in os.environ.get('KS', '256,512,1024,1184,1500,1776,2048').split(',')]:
    ctx.set_kernel_timing(True)
    ctx.kernel_times()
    for i in range(4):
        c, n, d = batches[i]
        ctx.extend(start, c[:K], n_ctrl=n[:K], dest=d, compare_f32=True, goals=goals)
    out[K] = round(float(ctx.kernel_times()[1:].mean()), 4)
print({k: os.environ.get(k) for k in ("MPS_ROLLOUT_CTAS_PER_SM", "MPS_ROLLOUT_GROUP", "MPS_ROLLOUT_MINB")}, out)
