#!/usr/bin/env python
"""bench.py — push rollouts/s (+ weighted-NN queries/s) of the B200 extend step.

Workload (BASELINE.json configs[1]): HybridActionRRT, robot + 6 objects (B = 7), one extend =
one nearest-neighbour query over the tree + K = 1024 chains of L <= 4 ramp controls propagated from
the selected start + argmin (float compare, pushing/algorithm/RRT.cpp:449-489).  A "step" is one
extend.  `value` = propagated controls ("push rollouts") per second with inputs resident in HBM;
`e2e` = the same through the host-buffer C ABI (mps_extend / mps_tree_nearest via the ctypes binding)
with the H2D of that step's controls + query and the D2H of the winner inside the timed region.
Multi-GPU (weak scaling): every rank propagates its own 1024-chain slice of a K = 1024 x N extend and
the winner records are all-gathered over NCCL.

The NN scan is additionally measured on BASELINE config 3's shape (B = 11, 1 M nodes) and reported
under "nn" with its own HBM roofline.

`--impl reference` times the CPU oracle (oracle/liboracle.so, a port: the reference's own path cannot
be built here) on all host threads on the same workload.
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

CFG = dict(n_objects=6, K=1024, L=4, tree_nodes=100_000, scene_seed=1002)
NN_CFG = dict(n_objects=10, nodes=1_000_000, seed=1003)
# BASELINE.json's metric names two figures; `value` is the first (push rollouts/sec), the second (weighted-NN
# queries/sec) is reported under "nn" of the same line
METRIC = "push rollouts/sec + weighted-NN queries/sec at 1/2/4/8 B200 vs host CPU"
UNIT = "rollouts/s"


def make_workload(rank: int, world: int, n_batches: int):
    from manipulation_planning_suite_b200 import scenes

    sc, _ = scenes.make_scene(CFG["n_objects"], seed=CFG["scene_seed"])
    start = scenes.cluttered_start(sc, CFG["scene_seed"])
    K, L = CFG["K"] * world, CFG["L"]
    batches = []
    for i in range(n_batches):
        rng = np.random.RandomState(7000 + i)
        controls = scenes.sample_controls_towards(rng, sc, start, K, L)
        n_ctrl = rng.randint(1, L + 1, size=K).astype(np.int32)
        dest = scenes.sample_state(rng, sc)
        batches.append((controls, n_ctrl, dest))
    goals = [(1, 0.5, 0.5, 0.1)]
    return sc, start, batches, goals


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        # NVML polled every few ms (the timed region is tens of ms long); nvidia-smi -lms as the fallback
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
    n_batches = steps + warmup
    sc, start, batches, goals = make_workload(rank, world, n_batches)
    B, S = sc.n_bodies, sc.n_statics
    K, L = CFG["K"], CFG["L"]

    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    with torch.cuda.stream(stream):
        ctx = api.Context(sc, device=local_rank, stream=stream.cuda_stream)
        tree = api.NearestNeighbors(ctx)
        tree.fill_random(CFG["tree_nodes"], seed=CFG["scene_seed"])
        sh = sharded.ShardedExtend(ctx)
        k0, k1 = sharded.shard_range(K * world, rank, world)

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident throughput (value) ---------------------------------------------
        # step i propagates batch i (the same batches the e2e leg uses); its inputs are uploaded before the
        # timed region of the step starts, the L2 is flushed, then the NN scan + extend are timed by events
        def rollouts_of(i):
            c, n, d = batches[i]
            r = ctx.extend(start, c[k0:k1], dump=True, n_ctrl=n[k0:k1], dest=d, compare_f32=True, goals=goals,
                           k_offset=k0)
            return int(((r.all_flags & 1) != 0).sum())

        per_batch = [rollouts_of(i) for i in range(n_batches)]
        for i in range(warmup):
            c, n, d = batches[i]
            sh.upload(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)
            tree.nearest_upload(d[None, :])
            tree.nearest_launch()
            sh.launch()
            flush.zero_()
        barrier()
        ctx.counters()  # reset
        launches0 = ctx.counters()["launches"]
        ctx.set_kernel_timing(True)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            c, n, d = batches[warmup + i]
            sh.upload(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)   # untimed: inputs resident
            tree.nearest_upload(d[None, :])
            flush.zero_()  # L2 flush before every timed step, outside the event pair
            ev[i][0].record(stream)
            tree.nearest_launch()
            sh.launch()
            ev[i][1].record(stream)
        barrier()
        clock_info = clocks.stop() if rank == 0 else None
        step_ms = np.array([a.elapsed_time(b) for a, b in ev], np.float64)
        total_ms = torch.tensor([step_ms.sum()], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        total_ms = float(total_ms.item())
        kt = ctx.kernel_times()
        ctx.set_kernel_timing(False)
        cnt = ctx.counters()
        launches = cnt["launches"] - launches0
        rollouts_local = sum(per_batch[warmup:warmup + steps])
        rollouts_all = torch.tensor([rollouts_local], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(rollouts_all, op=dist.ReduceOp.SUM)
        rollouts_per_step = float(rollouts_all.item()) / steps
        value = float(rollouts_all.item()) / (total_ms * 1e-3)
        winner = sh.result()

        # roofline of the dominant kernel (rollout_kernel): algorithmic flops / measured duration
        flops_per_launch = flop_model(B, S, sc.vel_iters, cnt) / max(steps, 1)
        k_ms = float(kt.mean()) if len(kt) else float("nan")
        fp32_peak = ctx.fp32_peak_tflops()
        achieved_tf = flops_per_launch / (k_ms * 1e-3) / 1e12
        roofline = {"kernel": "rollout_kernel", "bound": "fp32", "achieved": round(achieved_tf, 4),
                    "peak": round(fp32_peak, 2), "unit": "TFLOP/s", "frac": round(achieved_tf / fp32_peak, 5),
                    "traffic": ncu_traffic("rollout_kernel"), "kernel_ms": round(k_ms, 4), "share_of_step": round(k_ms * steps / step_ms.sum(), 4),
                    "peak_source": "measured in-run (FMA microbenchmark, csrc/peak.cu)",
                    "physics_steps_per_launch": cnt["steps"] / max(steps, 1),
                    "candidate_pairs_per_step": round(cnt["candidates"] / max(cnt["steps"], 1), 3),
                    "touching_pairs_per_step": round(cnt["touching"] / max(cnt["steps"], 1), 3),
                    "colour_passes_per_step": round(cnt["colour_passes"] / max(cnt["steps"], 1), 3),
                    "note": "latency/divergence-bound contact dynamics on CUDA cores; not a dense contraction"}

        # ---- end to end through the host-buffer C ABI (e2e) ------------------------------------
        h2d = d2h = 0
        for i in range(warmup):
            c, n, d = batches[i]
            tree.nearest(d)
            ctx.extend(start, c[k0:k1], n_ctrl=n[k0:k1], dest=d, compare_f32=True, goals=goals, k_offset=k0)
        barrier()
        t0 = time.perf_counter()
        t_nn = t_up = t_run = 0.0
        for i in range(warmup, warmup + steps):
            c, n, d = batches[i]
            ta = time.perf_counter()
            idx, dd = tree.nearest(d)                       # H2D query, D2H (index, distance)
            tb = time.perf_counter()
            sh.upload(start, c, n_ctrl=n, dest=d, compare_f32=True, goals=goals)   # H2D controls etc.
            tc = time.perf_counter()
            sh.launch()
            rec = sh.result()                               # D2H gathered winner records
            td = time.perf_counter()
            t_nn += tb - ta
            t_up += tc - tb
            t_run += td - tc
        barrier()
        t1 = time.perf_counter()
        e2e_t = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        # same batches as the device-resident leg
        e2e_local = sum(per_batch[warmup:warmup + steps])
        e2e_all = torch.tensor([e2e_local], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_all, op=dist.ReduceOp.SUM)
        e2e_value = float(e2e_all.item()) / float(e2e_t.item())
        h2d = 4 * (K * L * 5 + K + 3 * B * 2 + 4 * len(goals)) + 4 * 3 * B + 8
        d2h = sharded.record_size(B, L) * world + 16

    # ---- NN scan at BASELINE config 3's shape (rank 0 only) ----------------------------------
    nn = None
    if rank == 0 and not args.skip_nn:
        nn = bench_nn(api, scenes, local_rank)
        if world == 1 and not args.skip_cpu:
            nn["cpu_baseline"] = nn_cpu_baseline(scenes)

    throughput = None
    if rank == 0 and world == 1 and not args.skip_nn:
        throughput = bench_throughput(api, scenes, local_rank, fp32_peak)

    # ---- CPU baseline (oracle, 1 thread, bounded sample) on rank 0 at N = 1 -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_baseline(sc, start, batches[warmup:], goals, threads=1, max_seconds=12.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": round(total_ms / steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2 HybridActionRRT extend: B=7 (robot+6 objects), 4 wall boxes, "
                                   f"K={K} chains/GPU, L<=4 controls, tree {CFG['tree_nodes']} nodes, 1 NN query/extend",
                       "value_is": "push rollouts/sec (propagated controls per second); weighted-NN queries/sec is under nn",
                       "rollouts_per_step": rollouts_per_step, "l2": "flushed before every timed step (256 MB write)",
                       "sharding": f"K={K}x{world} chains sharded contiguously, winner records all-gathered (NCCL)"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(float(e2e_t.item()) / steps * 1e3, 4),
                    "host_ms_per_step": {"nn_query": round(t_nn / steps * 1e3, 4), "upload": round(t_up / steps * 1e3, 4),
                                         "launch_gather_fetch": round(t_run / steps * 1e3, 4)}},
            "gpu_launches": int(launches),
            "clocks": clock_info,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "nn": nn,
            "throughput": throughput,
            "winner": None if winner is None else {"k": winner.k, "n_ok": winner.n_ok, "distance": winner.distance},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_nn(api, scenes, device):
    """Single-query exact NN scans over 1 M nodes x 11 bodies (132 MB of states per query > 126 MB L2)."""
    import torch

    sc, _ = scenes.make_scene(NN_CFG["n_objects"], seed=NN_CFG["seed"])
    B = sc.n_bodies
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = api.Context(sc, device=device, stream=stream.cuda_stream)
        tree = api.NearestNeighbors(ctx)
        N = NN_CFG["nodes"]
        tree.fill_random(N, seed=NN_CFG["seed"], n_slices=64)
        rng = np.random.RandomState(11)
        q = scenes.sample_state(rng, sc)
        tree.nearest_upload(q[None, :])
        for _ in range(5):
            tree.nearest_launch()
        torch.cuda.synchronize()
        iters = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            tree.nearest_launch()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        idx, dist_ = tree.nearest_fetch()
        # end to end: host query in, (index, distance) out
        qs = [scenes.sample_state(rng, sc) for _ in range(53)]  # drawing the sample is not part of the query
        for i in range(3):
            tree.nearest(qs[i])
        t0 = time.perf_counter()
        for i in range(3, 53):
            tree.nearest(qs[i])
        e2e_ms = (time.perf_counter() - t0) / 50 * 1e3
        ctx.close()
    hbm, src = peaks()
    bytes_alg = N * 3 * B * 4
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    return {"workload": f"cfg3 shape: B={B}, N={N} nodes, full mask, 1 query per launch",
            "queries_per_s": round(1e3 / ms, 1), "us_per_query": round(ms * 1e3, 2),
            "e2e_queries_per_s": round(1e3 / e2e_ms, 1),
            "roofline": {"kernel": "nn_scan_kernel", "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm,
                         "unit": "GB/s", "frac": round(gbs / hbm, 4), "traffic": ncu_traffic("nn_scan_kernel"), "peak_source": src,
                         "algorithmic_bytes_per_launch": bytes_alg}}


def bench_throughput(api, scenes, device, fp32_peak):
    """The rollout kernel when the launch fills the machine (not the headline: BASELINE config 4's shape and
    the same with 7 bodies, single-control rollouts from a contact-rich start)."""
    import torch

    out = {}
    for name, n_objects, K in (("cfg4_shape_B17_K65536", 16, 65536), ("B7_K65536", 6, 65536)):
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


def cpu_baseline(sc, start, batches, goals, threads: int, max_seconds: float):
    """The CPU oracle (port of the path) timed on the host cores on a bounded sample of
    the same workload: whole batches of the timed steps, one after the other, until `max_seconds` of CPU
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
            "sample": f"{n} whole batches of {CFG['K']} chains from the timed steps ({done} propagated controls, {dt:.1f} s)"}


def run_reference(args):
    """Reference arm: the CPU oracle on all host threads (the reference's own C++ needs sim_env/Box2D/
    OMPL, which are not in this image: DESIGN.md §2), same workload, same metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ob = load_oracle()
    from manipulation_planning_suite_b200 import scenes  # noqa: F401  (scene synthesis only)

    steps, warmup = args.steps, args.warmup
    sc, start, batches, goals = make_workload(0, 1, steps + warmup)
    threads = os.cpu_count() or 1
    K = CFG["K"]
    # each step propagates the whole batch (1024 chains ~ 40 ms on 16 cores: no need to sub-sample)
    sample = K
    rng = np.random.RandomState(5)
    tree_states = np.stack([scenes.sample_state(rng, sc) for _ in range(20_000)])
    done = 0
    t_total = 0.0
    for i in range(steps + warmup):
        c, n, d = batches[i]
        t0 = time.perf_counter()
        ob.nn_linear(sc, tree_states, d)
        r = ob.extend(sc, start, c[:sample], n_ctrl=n[:sample], dest=d, compare_f32=True, goals=goals, threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            t_total += dt
            done += int(((r["all_flags"] & 1) != 0).sum())
    value = done / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warmup,
        "ms_per_step": round(t_total / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 HybridActionRRT extend: B=7 (robot+6 objects), 4 wall boxes, "
                               f"L<=4 controls; bounded sample of {sample} of the {K} chains per step, "
                               "20000-node linear NN scan per step"},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} chains per step x {steps} steps, pthread fan-out (one world per thread)"},
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
