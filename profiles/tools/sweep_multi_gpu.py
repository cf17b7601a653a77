"""BASELINE config 5 across the GPUs of one box: 4096 independent NaiveRRT instances (8 objects, K = 10),
partitioned contiguously over the ranks (512 per GPU at 8 GPUs), no communication until the final gather.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/sweep_multi_gpu.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, multi_seed, scenes  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n = total // world
sc, start = scenes.make_scene(8, seed=1005)
goals = [(1, float(start[3]) + 0.25, float(start[4]) + 0.1, 0.1)]
ctx = api.Context(sc, device=local)
# warm-up (context, kernels, allocator), then the timed sweep between barriers
DEVICE = os.environ.get("SWEEP_DEVICE", "0") == "1"
multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, num_control_samples=10, max_iterations=3, seed=99,
                           device_sampling=DEVICE)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res, forest = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, num_control_samples=10, max_iterations=iters,
                                         seed=5, goal_bias=0.2, target_bias=0.4, instance_offset=rank * n,
                                         device_sampling=DEVICE)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
v = torch.tensor([dt, float(res.rollouts), float((res.solved_iteration >= 0).sum())], dtype=torch.float64, device="cuda")
if world > 1:
    tmax = v[:1].clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(v[1:], op=dist.ReduceOp.SUM)
    v[0] = tmax[0]
print(f"rank {rank}: cpus {os.cpu_count()} affinity {len(os.sched_getaffinity(0))} wall {dt:.3f} timing {getattr(res, 'timing', None)}",
      file=sys.stderr, flush=True)
if rank == 0:
    print(json.dumps({"n_gpus": world, "instances": total, "instances_per_gpu": n, "iterations": iters,
                      "wall_s_max_over_ranks": round(float(v[0]), 3), "rollouts": int(v[1]),
                      "rollouts_per_s": round(float(v[1]) / float(v[0]), 1),
                      "solved_fraction": round(float(v[2]) / total, 4)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
