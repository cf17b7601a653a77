#!/usr/bin/env python
"""bench.py — push rollouts/s (+ weighted-NN queries/s) of the B200 extend step.

Headline workload (BASELINE.json configs[1]): HybridActionRRT, robot + 6 objects (B = 7).  One EXTEND = one
nearest-neighbour query over the tree + K = 1024 chains of L <= 4 ramp controls propagated from the selected start
+ argmin (float compare, pushing/algorithm/RRT.cpp:449-489).  A "step" of the contract is a burst of
EXTENDS_PER_STEP = 25 consecutive extends of fresh seeded batches, so that the default 20 timed steps cover 500
extends (about 1 s of device time) instead of 40 ms.  `value` = propagated controls ("push rollouts") per second
with every extend's inputs resident in HBM before its timed region starts (CUDA events per extend, L2 flushed
before each); `e2e` = the same through the host-buffer C ABI with the H2D of each extend's controls + query and
the D2H of the winner head inside the timed region.

Multi-GPU (`--gpus N`, one rank per GPU under torchrun), three curves per line:
  * weak (the headline): every rank propagates its own 1024-chain slice of a K = 1024 x N extend; slice s of
    batch i is the same controls at every N (seeded per slice); winner records all-gathered over NCCL and merged
    on the device;
  * "cfg4_strong": BASELINE configs[3] — the SAME 65 536 single-control rollouts (16-object clutter) split N ways;
  * "cfg5_sweep": BASELINE configs[4] — 4096 independent NaiveRRT instances (8 objects) / N per rank, run until
    90 % of them have reached the goal: time-to-solution distribution.

At N = 1 the line also carries the NN scan at configs[2]'s shape ("nn": L2 flushed between launches; exact top-k;
the two-level slice query), the rollout kernel in launches that fill the machine ("throughput"), and CPU context
for every reported shape ("cpu_context": the oracle on 1 thread and on all host threads, and the K below which
a CPU core is faster than the GPU call).

`--impl reference` times the CPU oracle (oracle/liboracle.so, a port: the reference's own path cannot be built
here) on all host threads on the same workload (same batches, same 100 000-node tree); it never
loads the CUDA library.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n_objects=6, K=1024, L=4, tree_nodes=100_000, scene_seed=1002, extends_per_step=25, batch_pool=32)
NN_CFG = dict(n_objects=10, nodes=1_000_000, seed=1003, n_slices=4096)
CFG4 = dict(n_objects=16, K=65536, seed=1016)
CFG5 = dict(n_objects=8, instances=4096, K=10, seed=1005, stop_fraction=0.9, max_iterations=1500)
# BASELINE.json's metric names two figures; `value` is the first (push rollouts/sec), the second (weighted-NN
# queries/sec) is reported under "nn" of the same line
METRIC = "push rollouts/sec + weighted-NN queries/sec at 1/2/4/8 B200 vs host CPU"
UNIT = "rollouts/s"
WORKLOAD = ("cfg2 HybridActionRRT extend: B=7 (robot+6 objects), 4 wall boxes, K={K} chains/GPU, L<=4 controls, "
            "tree {tree} nodes, 1 NN query/extend; step = {eps} consecutive extends of fresh batches")


def make_scene_and_start():
    from manipulation_planning_suite_b200 import scenes

    sc, _ = scenes.make_scene(CFG["n_objects"], seed=CFG["scene_seed"])
    start = scenes.cluttered_start(sc, CFG["scene_seed"])
    return sc, start


def make_slice(sc, start, batch: int, slice_id: int):
    """Chains [slice_id * K, (slice_id + 1) * K) of batch `batch`: the same controls whatever the number of ranks."""
    from manipulation_planning_suite_b200 import scenes

    rng = np.random.RandomState(7000 + batch + 1000 * slice_id)
    controls = scenes.sample_controls_towards(rng, sc, start, CFG["K"], CFG["L"])
    n_ctrl = rng.randint(1, CFG["L"] + 1, size=CFG["K"]).astype(np.int32)
    return controls, n_ctrl


def make_dest(sc, batch: int):
    from manipulation_planning_suite_b200 import scenes

    return scenes.sample_state(np.random.RandomState(9000 + batch), sc)


def make_workload(slice_ids, n_batches: int | None = None):
    """(scene, start, batches, goals); batches[i] = (controls [len(slice_ids) * K][L][5], n_ctrl, dest)"""
    sc, start = make_scene_and_start()
    n_batches = n_batches or CFG["batch_pool"]
    batches = []
    for i in range(n_batches):
        parts = [make_slice(sc, start, i, s) for s in slice_ids]
        batches.append((np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), make_dest(sc, i)))
    goals = [(1, 0.5, 0.5, 0.1)]
    return sc, start, batches, goals


def tree_states(sc, n: int):
    """The tree both arms scan: n seeded uniform states (host-generated so that the CPU arm sees the same nodes)."""
    rng = np.random.RandomState(CFG["scene_seed"])
    B = sc.n_bodies
    st = np.empty((n, 3 * B), np.float32)
    st[:, 0::3] = rng.uniform(sc.x_min, sc.x_max, size=(n, B))
    st[:, 1::3] = rng.uniform(sc.y_min, sc.y_max, size=(n, B))
    st[:, 2::3] = rng.uniform(-np.pi, np.pi, size=(n, B))
    return st


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe), NVML polled every 3 ms."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        self.nv = None
        self.samples = []
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, int(r)))
            except Exception:
                pass
            time.sleep(0.003)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            nv = self.nv
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted({k for _, _, r in self.samples for k, b in bits.items() if r & b})
            sm = [s_[0] for s_ in self.samples]
            mx = [s_[1] for s_ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def flop_model(B: int, S: int, vel_iters: int, counters: dict) -> float:
    """Algorithmic flops of the rollouts (SURVEY.md §8d counting model, I_p = 0 in mps-planar-v3):
    per step 10 (ramp) + 20 (B-1) (support friction) + 8 P (broad phase) + 30 B (integrate + pose),
    + 250 per candidate pair (narrow phase) + 2 * 55 * I_v per touching manifold (solver)."""
    P = B * (B - 1) // 2 + B * S
    per_step = 10 + 20 * (B - 1) + 8 * P + 30 * B
    return (counters["steps"] * per_step + 250.0 * counters["candidates"]
            + 2.0 * 55.0 * vel_iters * counters["touching"])


def ncu_traffic(kernel: str):
    """dram bytes per launch from the committed ncu capture (profiles/traffic.json), or None"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    from manipulation_planning_suite_b200 import api, scenes, sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the extend step has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps, warmup = args.steps, max(args.warmup, 3)
    EPS, POOL = CFG["extends_per_step"], CFG["batch_pool"]
    sc, start, batches, goals = make_workload([rank])   # this rank's slice of every batch of the pool
    B, S = sc.n_bodies, sc.n_statics
    K, L = CFG["K"], CFG["L"]
    k0 = rank * K
    legs = set(args.legs.split(","))

    def all_max(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    with torch.cuda.stream(stream):
        ctx = api.Context(sc, device=local_rank, stream=stream.cuda_stream)
        tree = api.NearestNeighbors(ctx)
        tree.add_batch(tree_states(sc, CFG["tree_nodes"]))
        sh = sharded.ShardedExtend(ctx)

        # propagated controls of this rank's slice of every batch of the pool (untimed, full dump)
        def rollouts_of(i):
            c, n, d = batches[i]
            r = ctx.extend(start, c, dump=True, n_ctrl=n, dest=d, compare_f32=True, goals=goals, k_offset=k0)
            return int(((r.all_flags & 1) != 0).sum())

        per_batch = [rollouts_of(i) for i in range(POOL)]

        def upload(i):
            c, n, d = batches[i % POOL]
            sh.upload_local(start, c, k0, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
            tree.nearest_upload(d[None, :])

        # ---- device-resident throughput (value) ---------------------------------------------
        # extend e propagates batch e % POOL; its inputs are uploaded before its timed region starts, the L2 is
        # flushed, then the NN scan + rollouts + select + gather + merge are timed by one event pair
        n_ext = steps * EPS
        for e in range(warmup * 2):
            upload(e)
            flush.zero_()
            tree.nearest_launch()
            sh.launch()
        barrier()
        ctx.counters()  # reset
        launches0 = ctx.counters()["launches"]
        ctx.set_kernel_timing(True)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_ext)]
        kt_all = []
        for e in range(n_ext):
            upload(warmup * 2 + e)      # untimed: inputs resident
            flush.zero_()               # L2 flush before every timed extend, outside the event pair
            ev[e][0].record(stream)
            tree.nearest_launch()
            sh.launch()
            ev[e][1].record(stream)
            if (e + 1) % 400 == 0:      # the library's per-launch timing ring holds 512 launches
                kt_all.append(ctx.kernel_times())
        barrier()
        clock_info = clocks.stop() if rank == 0 else None
        ext_ms = np.array([a.elapsed_time(b) for a, b in ev], np.float64)
        total_ms = all_max(float(ext_ms.sum()))
        kt_all.append(ctx.kernel_times())
        kt = np.concatenate(kt_all)
        ctx.set_kernel_timing(False)
        cnt = ctx.counters()
        launches = cnt["launches"] - launches0
        rollouts_local = sum(per_batch[(warmup * 2 + e) % POOL] for e in range(n_ext))
        rollouts_all = all_sum(float(rollouts_local))
        value = rollouts_all / (total_ms * 1e-3)
        head = sh.head()

        # roofline of the dominant kernel (rollout_kernel): algorithmic flops / measured duration
        flops_per_launch = flop_model(B, S, sc.vel_iters, cnt) / max(n_ext, 1)
        k_ms = float(kt.mean()) if len(kt) else float("nan")
        fp32_peak = ctx.fp32_peak_tflops()
        achieved_tf = flops_per_launch / (k_ms * 1e-3) / 1e12
        roofline = {"kernel": "rollout_kernel", "bound": "fp32", "achieved": round(achieved_tf, 4),
                    "peak": round(fp32_peak, 2), "unit": "TFLOP/s", "frac": round(achieved_tf / fp32_peak, 5),
                    "traffic": ncu_traffic("rollout_kernel"), "kernel_ms": round(k_ms, 4),
                    "kernel_ms_min_max": [round(float(kt.min()), 4), round(float(kt.max()), 4)] if len(kt) else None,
                    "launches_timed": int(len(kt)),
                    "share_of_step": round(k_ms * n_ext / ext_ms.sum(), 4),
                    "peak_source": "measured in-run (FMA microbenchmark, csrc/peak.cu)",
                    "physics_steps_per_launch": cnt["steps"] / max(n_ext, 1),
                    "candidate_pairs_per_step": round(cnt["candidates"] / max(cnt["steps"], 1), 3),
                    "touching_pairs_per_step": round(cnt["touching"] / max(cnt["steps"], 1), 3),
                    "colour_passes_per_step": round(cnt["colour_passes"] / max(cnt["steps"], 1), 3),
                    "note": "latency/divergence-bound contact dynamics on CUDA cores; not a dense contraction"}

        # ---- end to end through the host-buffer C ABI (e2e) ------------------------------------
        for e in range(warmup * 2):
            c, n, d = batches[e % POOL]
            tree.nearest(d)
            sh.upload_local(start, c, k0, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
            sh.launch()
            sh.head()
        barrier()
        t0 = time.perf_counter()
        t_nn = t_up = t_run = 0.0
        for e in range(n_ext):
            c, n, d = batches[(warmup * 2 + e) % POOL]
            ta = time.perf_counter()
            tree.nearest(d)                                 # H2D query (kernel parameters), D2H (index, distance)
            tb = time.perf_counter()
            sh.upload_local(start, c, k0, n_ctrl=n, dest=d, compare_f32=True, goals=goals)   # H2D controls etc.
            tc = time.perf_counter()
            sh.launch()                                     # rollouts, select, all-gather, merge: all on the device
            sh.head()                                       # D2H: the 24-byte head of the merged winner
            td = time.perf_counter()
            t_nn += tb - ta
            t_up += tc - tb
            t_run += td - tc
        barrier()
        e2e_s = all_max(time.perf_counter() - t0)
        e2e_value = rollouts_all / e2e_s
        h2d = EPS * (4 * (K * L * 5 + K + 3 * B * 2 + 4 * len(goals)) + 4 * 3 * B + 8)
        d2h = EPS * (24 + 16)

        # ---- the other two multi-GPU curves ----------------------------------------------------------
        cfg4 = bench_cfg4_strong(api, scenes, sharded, torch, stream, local_rank, rank, world, all_max, barrier, flush) \
            if "cfg4" in legs else None
    cfg5 = bench_cfg5_sweep(api, scenes, torch, dist, local_rank, rank, world, barrier) if "cfg5" in legs else None

    # ---- NN scan at BASELINE config 3's shape (rank 0, N = 1) ----------------------------------
    nn = None
    if rank == 0 and world == 1 and "nn" in legs:
        nn = bench_nn(api, scenes, local_rank, flush)
        if not args.skip_cpu:
            nn["cpu_baseline"] = nn_cpu_baseline(scenes)

    throughput = None
    if rank == 0 and world == 1 and "throughput" in legs:
        throughput = bench_throughput(api, scenes, local_rank, fp32_peak)

    # ---- CPU baseline (oracle, 1 thread, bounded sample) on rank 0 at N = 1 -------------------
    cpu = None
    cpu_ctx = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_baseline(sc, start, batches, goals, threads=1, max_seconds=10.0)
        if "cpu_context" in legs:
            cpu_ctx = cpu_context(api, scenes, sc, start, batches, goals, cpu)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": round(total_ms / steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(K=K, tree=CFG["tree_nodes"], eps=EPS),
                       "value_is": "push rollouts/sec (propagated controls per second); weighted-NN queries/sec is under nn",
                       "extends_per_step": EPS, "extends_timed": n_ext, "timed_region_s": round(total_ms * 1e-3, 3),
                       "ms_per_extend": round(total_ms / n_ext, 4),
                       "rollouts_per_extend": rollouts_all / n_ext,
                       "l2": "flushed before every timed extend (256 MB write, outside the event pair)",
                       "sharding": f"K={K}x{world}: rank r propagates slice r (the same controls at every N), winner "
                                   "records all-gathered (NCCL) and merged on the device"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_s / steps * 1e3, 4), "ms_per_extend": round(e2e_s / n_ext * 1e3, 4),
                    "host_ms_per_extend": {"nn_query": round(t_nn / n_ext * 1e3, 4), "upload": round(t_up / n_ext * 1e3, 4),
                                           "launch_gather_merge_head": round(t_run / n_ext * 1e3, 4)}},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "cpu_context": cpu_ctx,
            "nn": nn,
            "throughput": throughput,
            "cfg4_strong": cfg4,
            "cfg5_sweep": cfg5,
            "winner": {"k": head.k, "n_ok": head.n_ok, "distance": head.distance},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cfg4_workload(scenes):
    sc, _ = scenes.make_scene(CFG4["n_objects"], seed=CFG4["seed"])
    start = scenes.cluttered_start(sc, CFG4["seed"])
    rng = np.random.RandomState(CFG4["n_objects"])
    controls = scenes.sample_controls_towards(rng, sc, start, CFG4["K"], 1)
    dest = scenes.sample_state(rng, sc)
    return sc, start, controls, dest


def bench_cfg4_strong(api, scenes, sharded, torch, stream, device, rank, world, all_max, barrier, flush):
    """BASELINE configs[3]: ONE extend of 65 536 single-control rollouts from a 16-object clutter, the same rollouts
    at every N, chains [K r / N, K (r + 1) / N) on rank r, winner records all-gathered and merged on the device."""
    sc, start, controls, dest = cfg4_workload(scenes)
    K = CFG4["K"]
    k0, k1 = sharded.shard_range(K, rank, world)
    ctx = api.Context(sc, device=device, stream=stream.cuda_stream)
    sh = sharded.ShardedExtend(ctx)
    sh.upload_local(start, controls[k0:k1], k0, dest=dest)
    for _ in range(2):
        sh.launch()
    barrier()
    iters = 5
    ctx.set_kernel_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        flush.zero_()
        ev[i][0].record(stream)
        sh.launch()
        ev[i][1].record(stream)
    barrier()
    ms_local = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    ms = all_max(ms_local)
    k_local = float(ctx.kernel_times().mean())
    ctx.set_kernel_timing(False)
    per_rank = torch.tensor([ms_local, k_local], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        parts = [torch.empty_like(per_rank) for _ in range(world)]
        dist.all_gather(parts, per_rank)
        per_rank = torch.stack(parts)
    else:
        per_rank = per_rank[None, :]
    per_rank = per_rank.cpu().numpy()
    head = sh.head()
    # end to end: host controls in (this rank's slice), merged head out
    t0 = time.perf_counter()
    for i in range(3):
        sh.upload_local(start, controls[k0:k1], k0, dest=dest)
        sh.launch()
        sh.head()
    barrier()
    e2e_ms = all_max((time.perf_counter() - t0) / 3 * 1e3)
    ctx.close()
    return {"workload": f"cfg4 OracleRRT shape: B={sc.n_bodies}, the same K={K} single-control rollouts split over {world} GPU(s)",
            "scaling": "strong", "rollouts_per_s": round(K / ms * 1e3, 1), "ms_per_extend": round(ms, 4),
            "e2e_rollouts_per_s": round(K / e2e_ms * 1e3, 1), "e2e_ms_per_extend": round(e2e_ms, 4),
            "per_rank_ms_per_extend": [round(float(x), 4) for x in per_rank[:, 0]],
            "per_rank_rollout_kernel_ms": [round(float(x), 4) for x in per_rank[:, 1]],
            "h2d_bytes_per_extend_per_rank": int((k1 - k0) * 20 + 2 * 12 * sc.n_bodies), "d2h_bytes_per_extend": 24,
            "winner": {"k": head.k, "distance": head.distance}}


def cfg5_problem(scenes):
    sc, start = scenes.make_scene(CFG5["n_objects"], seed=CFG5["seed"])
    goals = [(1, float(start[3]) + 0.25, float(start[4]) + 0.1, 0.1)]
    return sc, start, goals


def bench_cfg5_sweep(api, scenes, torch, dist, device, rank, world, barrier):
    """BASELINE configs[4]: 4096 independent NaiveRRT instances (8 objects, K = 10), instance range split over the
    ranks, whole iterations on the device (mps_forest_naive_step), no communication until the final gather.  Every
    rank runs until 90 % of its instances have reached the goal (or max_iterations): time-to-solution."""
    from manipulation_planning_suite_b200 import multi_seed

    sc, start, goals = cfg5_problem(scenes)
    total = CFG5["instances"]
    n = total // world
    ctx = api.Context(sc, device=device)
    kw = dict(num_control_samples=CFG5["K"], goal_bias=0.2, target_bias=0.4, device_sampling=True)
    multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, max_iterations=3, seed=99, **kw)   # warm-up
    barrier()
    t0 = time.perf_counter()
    res, _ = multi_seed.naive_rrt_sweep(ctx, start, goals, n_instances=n, max_iterations=CFG5["max_iterations"], seed=5,
                                        instance_offset=rank * n, stop_fraction=CFG5["stop_fraction"], **kw)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = torch.from_numpy(np.where(np.isnan(res.solved_time), np.inf, res.solved_time)).cuda()
    agg = torch.tensor([wall, float(res.rollouts), float(res.iterations)], dtype=torch.float64, device="cuda")
    if world > 1:
        parts = [torch.empty_like(st) for _ in range(world)]
        dist.all_gather(parts, st)
        st = torch.cat(parts)
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        agg[0], agg[2] = mx[0], mx[2]
    ctx.close()
    t = np.sort(st.cpu().numpy())
    solved = np.isfinite(t)

    def at_fraction(f):   # seconds until the fraction f of ALL instances had a solution (inf = not reached)
        i = int(np.ceil(f * len(t))) - 1
        return None if not np.isfinite(t[i]) else round(float(t[i]), 4)

    return {"workload": f"cfg5: {total} NaiveRRT instances, B={sc.n_bodies}, K={CFG5['K']}, {n} per GPU, "
                        f"each rank stops at {int(CFG5['stop_fraction'] * 100)} % solved or {CFG5['max_iterations']} iterations",
            "scaling": "strong", "wall_s_max_over_ranks": round(float(agg[0]), 4),
            "solved_fraction": round(float(solved.mean()), 4), "iterations_max_over_ranks": int(agg[2]),
            "time_to_solution_s": {"first": round(float(t[0]), 4) if solved.any() else None, "p10": at_fraction(0.10),
                                   "p25": at_fraction(0.25), "p50": at_fraction(0.50), "p75": at_fraction(0.75),
                                   "p90": at_fraction(0.90)},
            "rollouts": int(agg[1]), "rollouts_per_s": round(float(agg[1]) / float(agg[0]), 1),
            "instances_solved_per_s": round(float(solved.sum()) / float(agg[0]), 1)}


def bench_nn(api, scenes, device, flush):
    """Exact NN scans over 1 M nodes x 11 bodies (132 MB of states per query > 126 MB L2), one query per launch,
    the L2 flushed before every timed launch (CUDA events per launch)."""
    import torch

    sc, _ = scenes.make_scene(NN_CFG["n_objects"], seed=NN_CFG["seed"])
    B = sc.n_bodies
    stream = torch.cuda.Stream()
    out = {}
    with torch.cuda.stream(stream):
        ctx = api.Context(sc, device=device, stream=stream.cuda_stream)
        tree = api.NearestNeighbors(ctx)
        N = NN_CFG["nodes"]
        tree.fill_random(N, seed=NN_CFG["seed"], n_slices=NN_CFG["n_slices"])
        rng = np.random.RandomState(11)
        q = scenes.sample_state(rng, sc)
        tree.nearest_upload(q[None, :])
        for _ in range(5):
            tree.nearest_launch()
        torch.cuda.synchronize()
        iters = 40
        # (a) the headline figure: three copies of the tree (3 x 136 MB) scanned in rotation, one query per launch -
        # by the time a tree is scanned again 272 MB of other rows have passed through the 126 MB L2, every launch
        # reads its 132 MB from HBM, and no dirty lines have to be written back while it runs
        others = []
        for j in range(2):
            c2 = api.Context(sc, device=device, stream=stream.cuda_stream)
            t2 = api.NearestNeighbors(c2)
            t2.fill_random(N, seed=NN_CFG["seed"], n_slices=NN_CFG["n_slices"])
            t2.nearest_upload(q[None, :])
            t2.nearest_launch()
            others.append((c2, t2))
        trees = [tree, others[0][1], others[1][1]]
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i in range(iters):
            ev[i][0].record(stream)
            trees[i % 3].nearest_launch()
            ev[i][1].record(stream)
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev[4:]]))
        # (a') the same rotation as a STREAM of scans under one event pair: the kernel's average launch duration when
        # queries are enqueued back to back.  Scans are launched with programmatic stream serialization; a scan that
        # follows another scan of its context with nothing in between starts streaming while that scan's last CTAs
        # finish (csrc/nn_scan.cu).  An event pair around a lone launch carries ~7 us that no kernel can remove (an
        # empty kernel of the same grid measures 7.0 us that way, profiles/tools/nn_trace_probe.cu).
        stream_ms = []
        for rep in range(5):
            trees[0].nearest_launch()
            e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0s.record(stream)
            for i in range(48):
                trees[(i + 1) % 3].nearest_launch()
            e1s.record(stream)
            torch.cuda.synchronize()
            stream_ms.append(e0s.elapsed_time(e1s) / 48)
        ms_stream = float(np.median(stream_ms))
        # the three metrics the planners query with (SURVEY.md 8d cfg 3): all bodies (RearrangementRRT::selectTreeNode),
        # all but the robot (ObjectArrangementDistanceFn), the robot alone (RobotStateDistanceFn); same rotation
        full = (1 << B) - 1
        per_mask = {}
        for name, m in (("all_bodies", full), ("all_but_robot", full & ~(1 << sc.robot_id)), ("robot_only", 1 << sc.robot_id)):
            for t in trees:
                t.nearest_upload(q[None, :], masks=np.array([m], np.uint32))
                t.nearest_launch()
            torch.cuda.synchronize()
            reps = []
            for rep in range(3):
                e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0s.record(stream)
                for i in range(48):
                    trees[(i + 1) % 3].nearest_launch()
                e1s.record(stream)
                torch.cuda.synchronize()
                reps.append(e0s.elapsed_time(e1s) / 48)
            t_ms = float(np.median(reps))
            nb = bin(m).count("1")
            per_mask[name] = {"us_per_query": round(t_ms * 1e3, 2), "queries_per_s": round(1e3 / t_ms, 1),
                              "GBps_algorithmic": round(N * 12 * nb / (t_ms * 1e-3) / 1e9, 1)}
        out["per_mask"] = per_mask
        tree.nearest_upload(q[None, :])
        tree.nearest_launch()
        for c2, t2 in others:
            c2.close()
        # (b) the L2 filled with DIRTY lines before every launch (256 MB memset): the scan then competes with the
        # write-back of up to 126 MB - harsher than a cold cache
        for i in range(iters):
            flush.zero_()
            ev[i][0].record(stream)
            tree.nearest_launch()
            ev[i][1].record(stream)
        torch.cuda.synchronize()
        ms_dirty = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        # back to back without a flush (what round 1 reported), for comparison
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            tree.nearest_launch()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_b2b = e0.elapsed_time(e1) / iters
        # end to end: host query in, (index, distance) out
        qs = [scenes.sample_state(rng, sc) for _ in range(103)]  # drawing the sample is not part of the query
        for i in range(3):
            tree.nearest(qs[i])
        t0 = time.perf_counter()
        for i in range(3, 103):
            tree.nearest(qs[i])
        e2e_ms = (time.perf_counter() - t0) / 100 * 1e3
        # the same 100 queries through the C entry point in a C loop (no Python / ctypes per call)
        _, _, sec = tree.nearest_timed(np.stack(qs[3:103]))
        e2e_c_ms = sec / 100 * 1e3
        # a batch of queries per launch (grid.y = Q): the per-launch fixed cost is shared
        Q = 32
        tree.nearest_upload(np.stack(qs[:Q]))
        tree.nearest_launch()
        torch.cuda.synchronize()
        flush.zero_()
        e0.record(stream)
        tree.nearest_launch()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_batch = e0.elapsed_time(e1)
        out["batch"] = {"queries_per_launch": Q, "us_per_query": round(ms_batch / Q * 1e3, 2),
                        "queries_per_s": round(Q / ms_batch * 1e3, 1),
                        "GBps_algorithmic": round(N * 3 * B * 4 * Q / (ms_batch * 1e-3) / 1e9, 1),
                        "note": "tiles of eight queries per CTA share every load (nn_scan_tile_kernel): arithmetic-bound, "
                                "the tree is read from HBM once per tile; GBps_algorithmic counts 132 MB per query"}
        # the same with 128 queries per launch (16 tiles: the grid is still one wave of the machine)
        Q2 = 128
        rng2 = np.random.RandomState(12)
        tree.nearest_upload(np.stack([scenes.sample_state(rng2, sc) for _ in range(Q2)]))
        tree.nearest_launch()
        torch.cuda.synchronize()
        flush.zero_()
        e0.record(stream)
        tree.nearest_launch()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_b2 = e0.elapsed_time(e1)
        out["batch_128"] = {"queries_per_launch": Q2, "us_per_query": round(ms_b2 / Q2 * 1e3, 2),
                            "queries_per_s": round(Q2 / ms_b2 * 1e3, 1)}
        if hasattr(tree, "nearest_k"):
            out["top_k"] = bench_nn_topk(tree, qs, flush, stream, torch, N, B)
        if hasattr(tree, "nearest_two_level"):
            out["slice_two_level"] = bench_nn_two_level(api, tree, qs, sc, flush, stream, torch, device)
        ctx.close()
    hbm, src = peaks()
    bytes_alg = N * 3 * B * 4
    gbs = bytes_alg / (ms_stream * 1e-3) / 1e9
    res = {"workload": f"cfg3 shape: B={B}, N={N} nodes, full mask, 1 query per launch, three trees scanned in rotation "
                       "(each launch reads its 132 MB from HBM); a stream of 48 launches under one CUDA event pair "
                       "(median of 5), and one event pair per launch beside it",
           "queries_per_s": round(1e3 / ms_stream, 1), "us_per_query": round(ms_stream * 1e3, 2),
           "us_per_query_lone_launch_event_pair": round(ms * 1e3, 2),
           "lone_launch_note": "an event pair around ONE launch includes ~7 us of launch and event latency (an empty "
                               "kernel of the same grid measures 7.0 us that way on this box class)",
           "us_per_query_after_256MB_memset": round(ms_dirty * 1e3, 2),
           "us_per_query_back_to_back_same_tree": round(ms_b2b * 1e3, 2),
           "e2e_queries_per_s": round(1e3 / e2e_c_ms, 1), "e2e_us_per_query": round(e2e_c_ms * 1e3, 2),
           "e2e_us_per_query_through_python_ctypes": round(e2e_ms * 1e3, 2),
           "e2e_note": "mps_tree_nearest with a host query and host results, one call per query, same tree every time "
                       "(a planner queries its one tree): the C entry point in a C loop, and the same through ctypes",
           "roofline": {"kernel": "nn_scan_kernel", "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm,
                        "unit": "GB/s", "frac": round(gbs / hbm, 4), "traffic": ncu_traffic("nn_scan_kernel"),
                        "peak_source": src, "algorithmic_bytes_per_launch": bytes_alg,
                        "frac_lone_launch": round(bytes_alg / (ms * 1e-3) / 1e9 / hbm, 4),
                        "note": "achieved = algorithmic bytes / average launch duration over the stream of launches; "
                                "the peak is a COPY bandwidth (read + write), a read-only stream can exceed it: "
                                "the nominal HBM3e figure is 7700 GB/s"}}
    res.update(out)
    return res


def bench_nn_topk(tree, qs, flush, stream, torch, N, B):
    """mps_tree_nearest_k: one streaming pass with warp-level top-k lists + exact fix-up (csrc/nn_scan.cu)"""
    res = {}
    for k in (1, 8, 32):
        tree.nearest_k(qs[0], k)
        ts = []
        for i in range(10):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tree.nearest_k(qs[1 + i], k)
            ts.append(time.perf_counter() - t0)
        ms = tree.ctx.time_launches(2, 10) / 10
        res[f"k{k}"] = {"e2e_us_per_query": round(float(np.median(ts)) * 1e6, 1), "kernel_us_back_to_back": round(ms * 1e3, 2),
                        "GBps_algorithmic": round(N * 3 * B * 4 / (ms * 1e-3) / 1e9, 1)}
    res["fallbacks"] = tree.nearest_k_fallbacks()
    return res


def bench_nn_two_level(api, tree, qs, sc, flush, stream, torch, device):
    """SliceBasedOracleRRT's query (RRT.cpp:783-827): nearest slice representative under the arrangement metric, then
    the nearest member of that slice under the robot-only metric.  Chained on the device (the second scan reads its
    slice filter from the first scan's result; one wait) against two host round trips."""
    B = sc.n_bodies
    rng = np.random.RandomState(3)
    n_slices = NN_CFG["n_slices"]
    rctx = api.Context(sc, device=device, stream=stream.cuda_stream)
    rtree = api.NearestNeighbors(rctx)
    reps = np.empty((n_slices, 3 * B), np.float32)
    reps[:, 0::3] = rng.uniform(sc.x_min, sc.x_max, size=(n_slices, B))
    reps[:, 1::3] = rng.uniform(sc.y_min, sc.y_max, size=(n_slices, B))
    reps[:, 2::3] = rng.uniform(-np.pi, np.pi, size=(n_slices, B))
    rtree.add_batch(reps)
    arr_mask, robot_mask = ((1 << B) - 1) & ~(1 << sc.robot_id), 1 << sc.robot_id
    out = {}
    for name in ("chained", "two_round_trips"):
        ts = []
        for i in range(23):
            if i >= 3:
                flush.zero_()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name == "chained":
                r = tree.nearest_two_level(rtree, qs[i], arr_mask, robot_mask)
            else:
                ri, rd = rtree.nearest(qs[i], mask=arr_mask)
                r = ((ri, rd), tree.nearest(qs[i], mask=robot_mask, slice_filter=ri))
            if i >= 3:
                ts.append(time.perf_counter() - t0)
        out[name + "_e2e_us_per_query"] = round(float(np.median(ts)) * 1e6, 1)
    out["n_slices"] = n_slices
    out["note"] = "second scan: 1 M nodes, robot sub-state rows only (12 MB) + the slice ids (4 MB)"
    rctx.close()
    return out


def bench_throughput(api, scenes, device, fp32_peak):
    """The rollout kernel when the launch fills the machine (not the headline: BASELINE config 4's shape, the
    same with 7 bodies, and config 5's per-iteration batch; single-control rollouts from a contact-rich start)."""
    out = {}
    for name, n_objects, K in (("cfg4_shape_B17_K65536", 16, 65536), ("B7_K65536", 6, 65536), ("cfg5_shape_B9_K40960", 8, 40960)):
        sc, _ = scenes.make_scene(n_objects, seed=1000 + n_objects)
        start = scenes.cluttered_start(sc, 1000 + n_objects)
        rng = np.random.RandomState(n_objects)
        controls = scenes.sample_controls_towards(rng, sc, start, K, 1)
        ctx = api.Context(sc, device=device)
        ctx.extend_upload(start, controls, dest=scenes.sample_state(rng, sc))
        ctx.extend_launch()
        ctx.synchronize()
        ctx.counters()
        iters = 3
        ms = ctx.time_launches(0, iters) / iters
        cnt = ctx.counters()
        flops = flop_model(sc.n_bodies, sc.n_statics, sc.vel_iters, cnt) / iters
        tf = flops / (ms * 1e-3) / 1e12
        out[name] = {"rollouts_per_s": round(K / ms * 1e3, 1), "ms_per_launch": round(ms, 3),
                     "physics_steps_per_launch": cnt["steps"] / iters,
                     "fp32": {"achieved_tflops": round(tf, 3), "peak_tflops": round(fp32_peak, 2),
                              "frac": round(tf / fp32_peak, 5)}}
        ctx.close()
    return out


def nn_cpu_baseline(scenes):
    """CPU side of the NN comparison on a bounded sample (200 k of the 1 M nodes): the oracle's exact linear
    scan and its restatement of ompl::NearestNeighborsGNAT (the reference's structure, RRT.cpp:46-51)."""
    ob = load_oracle()
    sc, _ = scenes.make_scene(NN_CFG["n_objects"], seed=NN_CFG["seed"])
    rng = np.random.RandomState(5)
    N = 200_000
    B = sc.n_bodies
    states = np.zeros((N, 3 * B), np.float32)
    states[:, 0::3] = rng.uniform(sc.x_min, sc.x_max, size=(N, B))
    states[:, 1::3] = rng.uniform(sc.y_min, sc.y_max, size=(N, B))
    states[:, 2::3] = rng.uniform(-np.pi, np.pi, size=(N, B))
    qs = [scenes.sample_state(rng, sc) for _ in range(8)]
    t0 = time.perf_counter()
    for q in qs:
        ob.nn_linear(sc, states, q)
    lin_us = (time.perf_counter() - t0) / len(qs) * 1e6
    g = ob.Gnat(sc)
    t0 = time.perf_counter()
    g.add_many(states)
    build_s = time.perf_counter() - t0
    e0 = g.distance_evaluations()
    t0 = time.perf_counter()
    for q in qs:
        g.nearest(q)
    gnat_us = (time.perf_counter() - t0) / len(qs) * 1e6
    frac = (g.distance_evaluations() - e0) / len(qs) / N
    return {"kind": "port", "cores": 1, "sample": f"{N} uniform nodes x {B} bodies, {len(qs)} queries",
            "linear_scan_us_per_query": round(lin_us, 1), "linear_scan_us_per_query_scaled_to_1M": round(lin_us * 5, 1),
            "gnat_us_per_query": round(gnat_us, 1), "gnat_build_s": round(build_s, 2),
            "gnat_distance_evaluations_per_node": round(frac, 3),
            "note": "on uniform states in SE(2)^11 the GNAT prunes almost nothing (distance concentration): "
                    "the reference's structure degenerates to a scan on this workload"}


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import oracle_binding as ob

    return ob


def cpu_rate(ob, sc, start, controls, n_ctrl, dest, goals, threads, max_seconds, chunk):
    """propagated controls per second of the oracle on `threads` threads: chunks of `chunk` chains of the given batch,
    one after the other, until max_seconds have passed"""
    t0 = time.perf_counter()
    done = chains = 0
    K = controls.shape[0]
    while time.perf_counter() - t0 < max_seconds:
        a = chains % K
        b = min(a + chunk, K)
        r = ob.extend(sc, start, controls[a:b], n_ctrl=None if n_ctrl is None else n_ctrl[a:b], dest=dest,
                      compare_f32=True, goals=goals, threads=threads)
        done += int(((r["all_flags"] & 1) != 0).sum())
        chains += b - a
    dt = time.perf_counter() - t0
    return done / dt, chains, dt


def cpu_baseline(sc, start, batches, goals, threads: int, max_seconds: float):
    """The CPU oracle (port of the path) timed on the host cores on a bounded sample of
    the same workload: whole batches of the timed extends, one after the other, until `max_seconds` of CPU
    work are done.  This is the only place bench.py executes oracle/ in the default arm."""
    ob = load_oracle()
    t0 = time.perf_counter()
    done = 0
    n = 0
    while time.perf_counter() - t0 < max_seconds:
        controls, n_ctrl, dest = batches[n % len(batches)]
        r = ob.extend(sc, start, controls, n_ctrl=n_ctrl, dest=dest, compare_f32=True, goals=goals, threads=threads)
        done += int(((r["all_flags"] & 1) != 0).sum())
        n += 1
    dt = time.perf_counter() - t0
    return {"value": round(done / dt, 1), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} whole batches of {CFG['K']} chains from the timed extends ({done} propagated controls, {dt:.1f} s)"}


def cpu_context(api, scenes, sc2, start2, batches2, goals2, cpu1):
    """CPU figures beside every reported shape (the oracle on 1 thread and on all host threads, nproc stated), and
    where the GPU call breaks even against the CPU for BASELINE configs[0]'s K = 10 extend."""
    ob = load_oracle()
    nproc = os.cpu_count() or 1
    out = {"nproc": nproc, "kind": "port (oracle/liboracle.so; pthread fan-out, one world per thread)"}
    c, n, d = batches2[0]
    r_all, chains, dt = cpu_rate(ob, sc2, start2, c, n, d, goals2, nproc, 2.5, 1024)
    out["cfg2"] = {"rollouts_per_s_1_thread": cpu1["value"], "rollouts_per_s_all_threads": round(r_all, 1),
                   "sample_all_threads": f"{chains} chains, {dt:.1f} s"}
    # cfg4 shape
    sc4, start4, controls4, dest4 = cfg4_workload(scenes)
    r1, ch1, dt1 = cpu_rate(ob, sc4, start4, controls4, None, dest4, None, 1, 2.5, 64)
    ra, cha, dta = cpu_rate(ob, sc4, start4, controls4, None, dest4, None, nproc, 2.5, 1024)
    out["cfg4_shape_B17"] = {"rollouts_per_s_1_thread": round(r1, 1), "rollouts_per_s_all_threads": round(ra, 1),
                             "sample": f"{ch1} / {cha} of the 65536 rollouts ({dt1:.1f} / {dta:.1f} s)"}
    # cfg5: the rollouts of NaiveRRT iterations (random controls from the problem's start state)
    sc5, start5, goals5 = cfg5_problem(scenes)
    rng = np.random.RandomState(5)
    controls5 = scenes.sample_controls(rng, 4096, 1)
    dest5 = scenes.sample_state(rng, sc5)
    r1, ch1, dt1 = cpu_rate(ob, sc5, start5, controls5, None, dest5, goals5, 1, 2.5, 64)
    ra, cha, dta = cpu_rate(ob, sc5, start5, controls5, None, dest5, goals5, nproc, 2.5, 1024)
    out["cfg5_B9"] = {"rollouts_per_s_1_thread": round(r1, 1), "rollouts_per_s_all_threads": round(ra, 1),
                      "sample": f"{ch1} / {cha} random single controls from the problem's start ({dt1:.1f} / {dta:.1f} s); the "
                                "GPU figure beside it is cfg5_sweep.rollouts_per_s (whole planner iterations)"}
    # cfg1: K = 10 extends (the reference's default num_control_samples, RRT.cpp:344), and the break-even K
    sc1, start1 = scenes.config_scene(1)
    controls1 = scenes.sample_controls(np.random.RandomState(1), 2048, 1)
    dest1 = scenes.sample_state(np.random.RandomState(2), sc1)
    r1, ch1, dt1 = cpu_rate(ob, sc1, start1, controls1, None, dest1, None, 1, 2.0, 10)
    ra, cha, dta = cpu_rate(ob, sc1, start1, controls1, None, dest1, None, nproc, 2.0, 10 * nproc)
    ctx = api.Context(sc1)
    gpu_ms = {}
    for K in (1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000):
        for _ in range(2):
            ctx.extend(start1, controls1[:K], dest=dest1)
        t0 = time.perf_counter()
        reps = 10
        for i in range(reps):
            ctx.extend(start1, controls1[:K], dest=dest1)
        gpu_ms[K] = (time.perf_counter() - t0) / reps * 1e3
    ctx.close()
    be1 = next((K for K in sorted(gpu_ms) if gpu_ms[K] < K / r1 * 1e3), None)
    bea = next((K for K in sorted(gpu_ms) if gpu_ms[K] < max(K / ra, 1.0 / r1) * 1e3), None)
    out["cfg1_B4"] = {"rollouts_per_s_1_thread": round(r1, 1), "rollouts_per_s_all_threads": round(ra, 1),
                      "cpu_ms_per_K10_extend_1_thread": round(10 / r1 * 1e3, 3),
                      "gpu_e2e_ms_per_extend_by_K": {str(k): round(v, 3) for k, v in gpu_ms.items()},
                      "gpu_rollouts_per_s_at_K10": round(10 / gpu_ms[10] * 1e3, 1),
                      "break_even_K_vs_1_thread": be1, "break_even_K_vs_all_threads": bea,
                      "note": "one chain costs the GPU the same latency as a thousand; below the break-even K a CPU core "
                              "finishes the extend sooner than the mps_extend call (host buffers in, winner out)"}
    return out


def run_reference(args):
    """Reference arm: the CPU oracle on all host threads (the reference's own C++ needs sim_env/Box2D/
    OMPL, which are not in this image: DESIGN.md §2), same workload, same metric, same tree.  Never loads the
    CUDA library (scene synthesis is pure Python)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ob = load_oracle()
    steps, warmup = args.steps, args.warmup
    EPS, POOL = CFG["extends_per_step"], CFG["batch_pool"]
    sc, start, batches, goals = make_workload([0])
    threads = os.cpu_count() or 1
    K = CFG["K"]
    tree = tree_states(sc, CFG["tree_nodes"])
    done = 0
    t_total = 0.0
    for s in range(steps + warmup):
        t0 = time.perf_counter()
        got = 0
        for e in range(EPS):
            # the same batch sequence as the GPU arm's timed extends for the timed steps
            c, n, d = batches[(max(warmup, 3) * 2 + (s - warmup) * EPS + e) % POOL]
            ob.nn_linear(sc, tree, d)
            r = ob.extend(sc, start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals, threads=threads)
            got += int(((r["all_flags"] & 1) != 0).sum())
        dt = time.perf_counter() - t0
        if s >= warmup:
            t_total += dt
            done += got
    value = done / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup,
        "ms_per_step": round(t_total / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(K=K, tree=CFG["tree_nodes"], eps=EPS),
                   "extends_per_step": EPS, "sample": "every step propagates its 25 whole batches (no sub-sampling); "
                                                      "the NN query is the oracle's exact linear scan"},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{K} chains per extend x {EPS} extends x {steps} steps, pthread fan-out (one world per thread)"},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_STDOUT_FD = None


def emit(line: dict):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def main():
    # libraries (NCCL, torch) may print banners on stdout: send everything but the JSON line to stderr
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-nn", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--legs", default="cfg4,cfg5,nn,throughput,cpu_context",
                    help="extra legs beside the headline (comma separated; empty string = headline only)")
    args = ap.parse_args()
    if args.skip_nn:
        args.legs = ",".join(x for x in args.legs.split(",") if x not in ("nn", "throughput"))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
