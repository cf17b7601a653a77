"""Multi-seed sweep: many independent NaiveRRT instances advanced in lockstep on one device
(BASELINE config 5: "4096 independent NaiveRRT instances ... time-to-solution").

Each instance is the reference's RearrangementRRT::plan loop with NaiveRearrangementRRT::extend
(pushing/algorithm/RRT.cpp:79-143, 171-204, 246-257, 358-390): sample (goal-biased or uniform valid
state) -> nearest tree node under the full metric -> K random controls propagated, best by the active
object's distance -> append.  All instances still running share one validity launch, one nearest-
neighbour launch and one extend launch per iteration (api.Forest); host code only draws random numbers
and keeps the bookkeeping.  Instances never interact: results depend on the seed, not on which other
instances are batched with them or on how many GPUs the sweep is split over.
"""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np

from . import abi, api, scenes


@dataclass
class SweepResult:
    solved_iteration: np.ndarray          # [n] iteration at which the instance reached the goal, -1 if not
    solved_time: np.ndarray               # [n] wall seconds at that moment, nan if not
    tree_sizes: np.ndarray                # [n]
    iterations: int
    rollouts: int                         # propagated controls over the whole sweep
    wall_s: float
    goal_nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))  # [n] node index of the goal state


def _sample_controls(rng: np.random.Generator, n: int) -> np.ndarray:
    """scenes.sample_controls (RampVelocityControlSampler::sample) in float32 on a Generator: [n][5]"""
    g = rng.standard_normal((n, 2), dtype=np.float32)
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
    u = rng.random((n, 3), dtype=np.float32)
    c = np.zeros((n, 5), np.float32)
    c[:, 0:2] = g * (scenes.V_LIMIT * np.sqrt(u[:, 0:1]))
    c[:, 2] = (2.0 * u[:, 1] - 1.0) * scenes.W_LIMIT
    c[:, 3] = scenes.DURATION_LIMITS[0] + (scenes.DURATION_LIMITS[1] - scenes.DURATION_LIMITS[0]) * u[:, 2]
    return c


def naive_rrt_sweep(ctx: api.Context, start: np.ndarray, goals, n_instances: int, num_control_samples: int = 10,
                    max_iterations: int = 500, seed: int = 1, goal_bias: float = 0.2, target_bias: float = 0.25,
                    robot_bias: float = 0.0, wall_timeout: float = 1e9, instance_offset: int = 0,
                    candidates_per_sample: int = 8, device_sampling: bool = False,
                    stop_fraction: float = 0.0) -> SweepResult:
    """stop_fraction > 0: the sweep ends as soon as that fraction of this call's instances has reached the goal
    (time-to-solution runs: the stragglers would otherwise run to max_iterations)."""
    sc = ctx.scene
    B = sc.n_bodies
    K = int(num_control_samples)
    if not (float(sc.x_max) - float(sc.x_min) < 3.0e38 and float(sc.y_max) - float(sc.y_min) < 3.0e38):
        raise ValueError("the scene's workspace bounds must span a finite range to sample states (default is +-FLT_MAX)")
    forest = api.Forest(ctx, n_instances, max_iterations + 2)
    forest.reset(np.tile(np.asarray(start, np.float32), (n_instances, 1)))
    targets = [int(g[0]) for g in goals]
    solved_it = np.full(n_instances, -1, np.int64)
    solved_t = np.full(n_instances, np.nan)
    goal_nodes = np.full(n_instances, -1, np.int64)
    active = np.arange(n_instances, dtype=np.int32)
    rollouts = 0
    timing = {"sample_host": 0.0, "is_valid": 0.0, "nearest": 0.0, "controls_host": 0.0, "extend": 0.0}
    t0 = time.perf_counter()
    it = 0
    C = candidates_per_sample
    T = len(targets)
    # Random numbers are drawn per BLOCK of 256 consecutive global instance ids from a generator seeded with
    # (seed, iteration, block): an instance's draws depend neither on which other instances are still active nor
    # on how the instances are split over GPUs (as long as the split is aligned to blocks), and a rank only
    # draws for its own instances.
    BLK = 256
    first_blk = instance_offset // BLK
    last_blk = (instance_offset + n_instances - 1) // BLK
    base_id = first_blk * BLK                      # global id of row 0 of the drawn arrays
    n_rows = (last_blk - first_blk + 1) * BLK

    def draw(iteration: int) -> dict:
        """All random numbers of one iteration for this rank's blocks.  Runs one iteration ahead on a worker
        thread while the device works (numpy and ctypes release the GIL)."""
        d = {"u_goal": np.empty(n_rows), "u_obj": np.empty(n_rows), "obj_uniform": np.empty(n_rows, np.int64),
             "tgt_pick": np.empty(n_rows, np.int64), "cand": np.empty((n_rows, C, B, 3), np.float32),
             "ball": np.empty((n_rows, C, T, 2), np.float32), "controls": np.empty((n_rows, K, 1, 5), np.float32)}
        for blk in range(first_blk, last_blk + 1):
            rng = np.random.Generator(np.random.SFC64([int(seed), int(iteration), int(blk)]))
            r0 = (blk - first_blk) * BLK
            sl = slice(r0, r0 + BLK)
            d["u_goal"][sl] = rng.random(BLK)
            d["u_obj"][sl] = rng.random(BLK)
            d["obj_uniform"][sl] = rng.integers(0, B, size=BLK)
            d["tgt_pick"][sl] = np.minimum(rng.integers(0, T + 1, size=BLK), T - 1)  # :133-136 quirk
            d["cand"][sl] = rng.random((BLK, C, B, 3), dtype=np.float32)
            ball = rng.standard_normal((BLK, C, T, 2), dtype=np.float32)
            ball /= np.maximum(np.linalg.norm(ball, axis=-1, keepdims=True), 1e-12)
            ball *= np.sqrt(rng.random((BLK, C, T, 1), dtype=np.float32))
            d["ball"][sl] = ball
            d["controls"][sl] = _sample_controls(rng, BLK * K).reshape(BLK, K, 1, 5)
        return d

    if device_sampling:
        # the whole iteration runs on the device (mps_forest_naive_step): no host random numbers, no uploads
        # but the ids of the instances still running; draws are a hash of (seed, iteration, global instance id)
        for it in range(max_iterations):
            if len(active) == 0 or time.perf_counter() - t0 > wall_timeout:
                break
            if stop_fraction > 0.0 and n_instances - len(active) >= stop_fraction * n_instances:
                break
            ta = time.perf_counter()
            res = forest.naive_step(active, it, seed, goals, K=K, C_=C, goal_bias=goal_bias, target_bias=target_bias,
                                    robot_bias=robot_bias, instance_offset=instance_offset)
            timing["extend"] += time.perf_counter() - ta
            rollouts += len(active) * K
            reached = ((res["flags"] & abi.MPS_F_GOAL) != 0) & (res["k"] >= 0)
            if reached.any():
                ids = active[reached]
                solved_it[ids] = it
                solved_t[ids] = time.perf_counter() - t0
                goal_nodes[ids] = res["tree_index"][reached]
                active = active[~reached]
        wall = time.perf_counter() - t0
        res_ = SweepResult(solved_it, solved_t, forest.sizes(), it + 1, rollouts, wall, goal_nodes)
        res_.timing = {"device_step": round(timing["extend"], 4)}
        return res_, forest
    pool = ThreadPoolExecutor(max_workers=1)
    ahead = pool.submit(draw, 0)
    lo = np.array([sc.x_min, sc.y_min, -np.pi], np.float32)
    span = np.array([sc.x_max - sc.x_min, sc.y_max - sc.y_min, 2 * np.pi], np.float32)
    for it in range(max_iterations):
        if len(active) == 0 or time.perf_counter() - t0 > wall_timeout:
            break
        if stop_fraction > 0.0 and n_instances - len(active) >= stop_fraction * n_instances:
            break
        ta = time.perf_counter()
        rnd = ahead.result()
        ahead = pool.submit(draw, it + 1)
        gid = active.astype(np.int64) + (instance_offset - base_id)  # row of each active instance in the drawn arrays
        u_goal = rnd["u_goal"][gid]
        u_obj = rnd["u_obj"][gid]
        obj_uniform = rnd["obj_uniform"][gid]
        tgt_pick = rnd["tgt_pick"][gid]
        M = len(active)
        # ---- sample(): goal-biased or uniform VALID state; C candidates each, one validity launch ----
        cand = (lo + span * rnd["cand"][gid]).reshape(M, C, 3 * B)
        ball = rnd["ball"][gid]
        is_goal = u_goal < goal_bias
        for gi, g in enumerate(goals):  # ObjectsRelocationGoal::sampleGoal: targets inside their goal discs
            b = int(g[0])
            cand[is_goal, :, 3 * b] = g[1] + g[3] * ball[is_goal, :, gi, 0]
            cand[is_goal, :, 3 * b + 1] = g[2] + g[3] * ball[is_goal, :, gi, 1]
        tb = time.perf_counter()
        valid, _ = ctx.is_valid(cand.reshape(M * C, 3 * B))
        tc = time.perf_counter()
        valid = valid.reshape(M, C)
        first = np.where(valid.any(axis=1), valid.argmax(axis=1), C - 1)
        samples = cand[np.arange(M), first]
        # active object: target of the goal, else sampleActiveObject (RRT.cpp:246-257)
        obj = np.where(u_obj < target_bias, np.asarray(targets)[tgt_pick],
                       np.where(u_obj < target_bias + robot_bias, sc.robot_id, obj_uniform))
        obj = np.where(is_goal, np.asarray(targets)[tgt_pick], obj).astype(np.int64)
        # ---- selectTreeNode(): nearest node under the full metric (RRT.cpp:192-204) ----
        td = time.perf_counter()
        near, _ = forest.nearest(active, samples)
        te = time.perf_counter()
        # ---- extend(): K random controls, best by the active object's distance, appended (:358-390) ----
        controls = rnd["controls"][gid]
        tf = time.perf_counter()
        res = forest.extend(active, near, controls, dests=samples, masks=(1 << obj).astype(np.uint32), goals=goals,
                            append=True)
        tg = time.perf_counter()
        timing["sample_host"] += (tb - ta) + (td - tc)
        timing["is_valid"] += tc - tb
        timing["nearest"] += te - td
        timing["controls_host"] += tf - te
        timing["extend"] += tg - tf
        rollouts += M * K
        reached = (res["flags"][:, 0] & abi.MPS_F_GOAL) != 0
        reached &= res["k"] >= 0
        if reached.any():
            now = time.perf_counter() - t0
            ids = active[reached]
            solved_it[ids] = it
            solved_t[ids] = now
            goal_nodes[ids] = res["tree_index"][reached]
            active = active[~reached]
    wall = time.perf_counter() - t0
    ahead.cancel()
    pool.shutdown(wait=True)
    res_ = SweepResult(solved_it, solved_t, forest.sizes(), it + 1, rollouts, wall, goal_nodes)
    res_.timing = {k: round(v, 4) for k, v in timing.items()}
    return res_, forest


def extract_path(forest: api.Forest, inst: int, goal_node: int):
    """root .. goal_node of one instance: (states [n][3B], controls [n][5])"""
    n = goal_node + 1
    st, pa, co = forest.read(inst, 0, n)
    idx = []
    cur = goal_node
    while cur >= 0:
        idx.append(cur)
        cur = int(pa[cur])
    idx = idx[::-1]
    return st[idx], co[idx]
