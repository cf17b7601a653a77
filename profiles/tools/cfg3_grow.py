"""BASELINE config 3 as specified: a tree GROWN by SliceOracleRRT itself (RRT.cpp:766-884) on the 11-body scene,
not filled with seeded states.  The goal is out of reach (the target would have to leave the workspace), so the
planner keeps growing its tree until the iteration budget or the time-out; every iteration is the reference's:
sample -> nearest slice representative under the arrangement metric -> nearest member of that slice under the
robot metric -> oracle extend -> append (new slice when the arrangement changed).

Reports per budget: nodes, slices, iterations / s, NN queries / s, propagations / s.  The cost of a query grows with
the tree (the scan is exhaustive): the last column is the time one single query takes over the grown tree's size.

usage: python profiles/tools/cfg3_grow.py [iterations ...]     (default 20000 100000)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manipulation_planning_suite_b200 import api, host_api, scenes  # noqa: E402

budgets = [int(a) for a in sys.argv[1:]] or [20000, 100000]
sc, start = scenes.config_scene(3)
# object 1 would have to be pushed 10 m beyond the walls: never satisfied, never sampled as a goal (goal_bias 0)
goals = [(1, float(sc.x_max) + 10.0, float(sc.y_max) + 10.0, 0.05)]
for n_it in budgets:
    t0 = time.perf_counter()
    r = host_api.plan(sc, start, goals, algorithm="SliceOracleRRT", seed=3, time_out=3600.0, max_iterations=n_it,
                      num_control_samples=10, goal_bias=0.0, target_bias=0.1, robot_bias=0.1)
    wall = time.perf_counter() - t0
    s = r.stats
    # what one query over a tree of this size costs (seeded states: the scan does not care what the nodes hold)
    ctx = api.Context(sc)
    tree = api.NearestNeighbors(ctx)
    tree.fill_random(max(int(s["tree_size"]), 1), seed=7, n_slices=max(int(s["num_slices"]), 1))
    qs = np.stack([scenes.sample_state(np.random.RandomState(i), sc) for i in range(101)])
    _, _, sec = tree.nearest_timed(qs)
    ctx.close()
    print(json.dumps({"iterations": s["num_iterations"], "tree_nodes": s["tree_size"], "slices": s["num_slices"],
                      "wall_s": round(wall, 2), "iterations_per_s": round(s["num_iterations"] / wall, 1),
                      "nn_queries": s["num_nearest_neighbor_queries"],
                      "nn_queries_per_s": round(s["num_nearest_neighbor_queries"] / wall, 1),
                      "propagations": s["num_state_propagations"],
                      "propagations_per_s": round(s["num_state_propagations"] / wall, 1),
                      "single_query_us_at_this_size": round(sec / 101 * 1e6, 2), "solved": bool(r.solved)}), flush=True)
