"""Thin Python binding over the C ABI (include/mps_b200.h) used by tests, bench.py and the
multi-GPU driver.  Class and method names follow the reference's seams:

  Context.propagate      <- SimEnvStatePropagator::propagate   (ompl/control/SimEnvStatePropagator.cpp:42-118)
  Context.extend         <- NaiveControlSampler::sampleTo / HybridActionRRT::extend K-loop
                            (ompl/control/NaiveControlSampler.cpp:34-71, pushing/algorithm/RRT.cpp:449-489)
  Context.is_valid       <- SimEnvValidityChecker::isValid     (ompl/state/SimEnvState.cpp:1441-1462)
  NearestNeighbors       <- ompl::NearestNeighbors<MotionPtr> as used by RearrangementRRT
                            (pushing/algorithm/RRT.cpp:46-51,202,239): add / nearest / clear / size

Every call goes to libmps_b200.so; nothing here computes.  Errors raise MpsError with the
library's message; a missing library or device is a hard error (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi


class MpsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mps error {code}: {msg}")
        self.code = code


def _fp(a: np.ndarray | None, t=C.c_float):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a, shape=None) -> np.ndarray | None:
    if a is None:
        return None
    out = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        out = out.reshape(shape)
    return out


@dataclass
class ExtendResult:
    k: int               # winning rollout (global index) or -1
    n_ok: int
    goal_step: int
    tree_index: int
    distance: float
    best_states: np.ndarray    # [L, 3B]
    best_controls: np.ndarray  # [L, 5]
    best_flags: np.ndarray     # [L]
    all_states: np.ndarray | None = None   # [K, L, 3B]
    all_rest: np.ndarray | None = None     # [K, L]
    all_flags: np.ndarray | None = None    # [K, L]
    all_dist: np.ndarray | None = None     # [K]
    all_steps: np.ndarray | None = None    # [K, L]


class Context:
    """One planner instance on one device (struct mps_ctx)."""

    def __init__(self, scene: abi.Scene, device: int = 0, stream: int | None = None):
        self.lib = abi.load_library()
        self.scene = scene
        self.B = int(scene.n_bodies)
        h = C.c_void_p()
        if stream is None:
            rc = self.lib.mps_ctx_create(C.byref(scene), device, C.byref(h))
        else:
            rc = self.lib.mps_ctx_create_on_stream(C.byref(scene), device, C.c_void_p(stream), C.byref(h))
        if rc != abi.MPS_OK:
            raise MpsError(rc, (self.lib.mps_last_error(None) or b"").decode())
        self.h = h
        self.device = int(device)
        self._keep = None  # arrays referenced by an uploaded extend

    def close(self):
        if getattr(self, "h", None):
            self.lib.mps_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != abi.MPS_OK:
            raise MpsError(rc, (self.lib.mps_last_error(self.h) or b"").decode())

    def synchronize(self):
        self._check(self.lib.mps_ctx_synchronize(self.h))

    # ---- propagation ----------------------------------------------------
    def _extend_in(self, start, controls, n_ctrl=None, dest=None, weights=None, mask=None, compare_f32=False,
                   goals=None, k_offset=0, ball_center=None, ball_radius=0.0, append_to_tree=False, parent=-1,
                   starts=None, stop_at_goal=False):
        B = self.B
        controls = _f32(controls)
        if controls.ndim == 2:
            controls = controls[:, None, :]
        K, L, five = controls.shape
        assert five == abi.MPS_CONTROL_DIM
        start = None if start is None else _f32(start, (3 * B,))
        dest = _f32(dest, (3 * B,))
        weights = _f32(weights, (B,))
        ball_center = _f32(ball_center, (3 * B,))
        n_ctrl_a = None if n_ctrl is None else np.ascontiguousarray(n_ctrl, dtype=np.int32).reshape(K)
        garr = None
        ng = 0
        if goals:
            ng = len(goals)
            garr = (abi.Goal * ng)()
            for i, g in enumerate(goals):
                garr[i].body, garr[i].x, garr[i].y, garr[i].tol = int(g[0]), float(g[1]), float(g[2]), float(g[3])
        ein = abi.ExtendIn()
        ein.start = _fp(start)
        ein.controls = _fp(controls)
        ein.n_ctrl = _fp(n_ctrl_a, C.c_int32)
        ein.K, ein.L = K, L
        ein.dest = _fp(dest)
        ein.weights = _fp(weights)
        full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
        ein.mask = full if mask is None else int(mask)
        ein.compare_f32 = 1 if compare_f32 else 0
        ein.goals = garr if garr is not None else None
        ein.n_goals = ng
        ein.k_offset = int(k_offset)
        ein.ball_center = _fp(ball_center)
        ein.ball_radius = float(ball_radius)
        ein.append_to_tree = 1 if append_to_tree else 0
        ein.parent = int(parent)
        starts_a = None if starts is None else _f32(starts).reshape(K, 3 * B)
        ein.starts = _fp(starts_a)
        ein.stop_at_goal = 1 if stop_at_goal else 0
        keep = (start, controls, n_ctrl_a, dest, weights, ball_center, garr, starts_a)
        return ein, keep, K, L

    def _extend_out(self, K, L, dump):
        B = self.B
        best = abi.ExtendBest()
        res = dict(
            best=best,
            best_states=np.zeros((L, 3 * B), np.float32),
            best_controls=np.zeros((L, 5), np.float32),
            best_flags=np.zeros(L, np.uint32),
        )
        eout = abi.ExtendOut()
        eout.best = C.pointer(best)
        eout.best_states = _fp(res["best_states"])
        eout.best_controls = _fp(res["best_controls"])
        eout.best_flags = _fp(res["best_flags"], C.c_uint32)
        if dump:
            res["all_states"] = np.zeros((K, L, 3 * B), np.float32)
            res["all_rest"] = np.zeros((K, L), np.float32)
            res["all_flags"] = np.zeros((K, L), np.uint32)
            res["all_dist"] = np.zeros(K, np.float64)
            res["all_steps"] = np.zeros((K, L), np.int32)
            eout.all_states = _fp(res["all_states"])
            eout.all_rest = _fp(res["all_rest"])
            eout.all_flags = _fp(res["all_flags"], C.c_uint32)
            eout.all_dist = _fp(res["all_dist"], C.c_double)
            eout.all_steps = _fp(res["all_steps"], C.c_int32)
        return eout, res

    @staticmethod
    def _result(res) -> ExtendResult:
        b = res["best"]
        return ExtendResult(
            k=b.k, n_ok=b.n_ok, goal_step=b.goal_step, tree_index=b.tree_index, distance=b.distance,
            best_states=res["best_states"], best_controls=res["best_controls"], best_flags=res["best_flags"],
            all_states=res.get("all_states"), all_rest=res.get("all_rest"), all_flags=res.get("all_flags"),
            all_dist=res.get("all_dist"), all_steps=res.get("all_steps"),
        )

    def extend(self, start, controls, dump=False, **kw) -> ExtendResult:
        """K rollout chains from `start`, best-of by distance to `dest` (host buffers, synchronous)."""
        ein, keep, K, L = self._extend_in(start, controls, **kw)
        eout, res = self._extend_out(K, L, dump)
        self._check(self.lib.mps_extend(self.h, C.byref(ein), C.byref(eout)))
        del keep
        return self._result(res)

    def extend_upload(self, start, controls, **kw):
        ein, keep, K, L = self._extend_in(start, controls, **kw)
        self._check(self.lib.mps_extend_upload(self.h, C.byref(ein)))
        self._ext_shape = (K, L)

    def extend_launch(self):
        self._check(self.lib.mps_extend_launch(self.h))

    @staticmethod
    def hybrid_gen(scene: abi.Scene, seed: int, iteration: int = 0, k_offset: int = 0, fixed_object: int = -1,
                   action_randomness: float = 0.0, target_bias: float = 0.0, robot_bias: float = 0.0, targets=(),
                   velocity_limits=(0.6, 0.6, 1.5), duration_limits=(0.1, 1.0), optimal_push_distance: float = 0.2,
                   push_distance_tolerance: float = 0.04, push_angle_tolerance: float = 0.03, object_size=None) -> abi.HybridGen:
        """struct mps_hybrid_gen with the reference's defaults (HumanOracle.cpp:10-18, OraclePushPlanner.cpp:70-71);
        object_size defaults to max(width, height) of the scene's bodies"""
        g = abi.HybridGen()
        g.seed, g.iteration, g.k_offset, g.fixed_object = int(seed), int(iteration), int(k_offset), int(fixed_object)
        g.action_randomness, g.target_bias, g.robot_bias = action_randomness, target_bias, robot_bias
        g.n_targets = len(targets)
        for i, t in enumerate(targets):
            g.targets[i] = int(t)
        for i in range(3):
            g.velocity_limits[i] = velocity_limits[i]
        g.duration_min, g.duration_max = duration_limits
        g.optimal_push_distance, g.push_distance_tolerance = optimal_push_distance, push_distance_tolerance
        g.push_angle_tolerance = push_angle_tolerance
        for i in range(int(scene.n_bodies)):
            if object_size is not None:
                g.object_size[i] = object_size[i]
            elif scene.shape[i] == abi.MPS_SHAPE_DISC:
                g.object_size[i] = 2.0 * scene.hx[i]
            else:
                g.object_size[i] = 2.0 * max(scene.hx[i], scene.hy[i])
        return g

    def extend_upload_generated(self, start, K: int, L: int, gen: abi.HybridGen, **kw):
        """mps_extend_upload_generated: the K chains' action sequences are produced on the device from `gen`"""
        dummy = np.zeros((1, 1, 5), np.float32)
        ein, keep, _, _ = self._extend_in(start, dummy, **kw)
        ein.controls = None
        ein.n_ctrl = None
        ein.K, ein.L = int(K), int(L)
        self._check(self.lib.mps_extend_upload_generated(self.h, C.byref(ein), C.byref(gen)))
        self._ext_shape = (int(K), int(L))

    def generated_controls(self):
        """(controls [K][L][5], n_ctrl [K], n_truncated) of the last extend_upload_generated"""
        K, L = self._ext_shape
        c = np.zeros((K, L, 5), np.float32)
        n = np.zeros(K, np.int32)
        tr = C.c_int32(0)
        self._check(self.lib.mps_extend_generated_controls(self.h, _fp(c), _fp(n, C.c_int32), C.byref(tr)))
        return c, n, tr.value

    def sample_controls(self, seed: int, n: int, iteration: int = 0, gid: int = 0, velocity_limit: float = 0.6,
                        omega_limit: float = 1.5, duration_limits=(0.1, 1.0)) -> np.ndarray:
        c = np.zeros((n, 5), np.float32)
        self._check(self.lib.mps_sample_controls(self.h, int(seed), int(iteration), int(gid), int(n), velocity_limit, omega_limit,
                                                 duration_limits[0], duration_limits[1], _fp(c)))
        return c

    def extend_fetch(self, dump=False) -> ExtendResult:
        K, L = self._ext_shape
        eout, res = self._extend_out(K, L, dump)
        self._check(self.lib.mps_extend_fetch(self.h, C.byref(eout)))
        return self._result(res)

    def propagate(self, state, control):
        """bool propagate(state, control /*rest time mutated*/, result). Returns (ok, result, control, flags)."""
        B = self.B
        s = _f32(state, (3 * B,)).copy()
        c = _f32(control, (5,)).copy()
        out = np.zeros(3 * B, np.float32)
        ok = C.c_int32(0)
        fl = C.c_uint32(0)
        self._check(self.lib.mps_propagate(self.h, _fp(s), _fp(c), _fp(out), C.byref(ok), C.byref(fl)))
        return bool(ok.value), out, c, int(fl.value)

    def counters(self) -> dict:
        arr = (C.c_uint64 * 8)()
        self._check(self.lib.mps_extend_counters(self.h, arr))
        return dict(steps=arr[0], candidates=arr[1], touching=arr[2], colour_passes=arr[3], launches=arr[4])

    def chain_cycles(self) -> np.ndarray:
        K = self._ext_shape[0]
        out = np.zeros(K, np.int64)
        self._check(self.lib.mps_extend_chain_cycles(self.h, _fp(out, C.c_int64)))
        return out

    def time_launches(self, what: int, iters: int) -> float:
        ms = C.c_float(0)
        self._check(self.lib.mps_time_launches(self.h, what, iters, C.byref(ms)))
        return float(ms.value)

    def set_kernel_timing(self, enabled: bool):
        self._check(self.lib.mps_ctx_set_kernel_timing(self.h, 1 if enabled else 0))

    def kernel_times(self) -> np.ndarray:
        buf = np.zeros(512, np.float32)
        n = C.c_int32(0)
        self._check(self.lib.mps_extend_kernel_times(self.h, _fp(buf), 512, C.byref(n)))
        return buf[: n.value].copy()

    def fp32_peak_tflops(self) -> float:
        v = C.c_float(0)
        self._check(self.lib.mps_measure_fp32_peak(self.h, C.byref(v)))
        return float(v.value)

    def record_size(self) -> int:
        n = C.c_int64(0)
        self._check(self.lib.mps_extend_record_size(self.h, C.byref(n)))
        return n.value

    def record_to(self, device_ptr: int):
        self._check(self.lib.mps_extend_record_to(self.h, C.c_void_p(device_ptr)))

    def merge_records(self, gathered_ptr: int, world: int, has_dest: bool, compare_f32: bool, append_to_tree: bool = False,
                      parent: int = -1, wait: bool = False):
        """Exchange step of the K-sharded extend on the device (mps_extend_merge_records).  wait=True returns the
        24-byte head (abi.ExtendBest), else the call is asynchronous on the ctx stream."""
        best = abi.ExtendBest() if wait else None
        self._check(self.lib.mps_extend_merge_records(self.h, C.c_void_p(gathered_ptr), world, int(has_dest), int(compare_f32),
                                                      int(append_to_tree), parent, C.byref(best) if wait else None))
        return best

    def extend_head(self) -> abi.ExtendBest:
        best = abi.ExtendBest()
        self._check(self.lib.mps_extend_head(self.h, C.byref(best)))
        return best

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.mps_ctx_stream(self.h, C.byref(p)))
        return int(p.value or 0)

    # ---- validity -------------------------------------------------------
    def is_valid(self, states):
        st = _f32(states).reshape(-1, 3 * self.B)
        n = st.shape[0]
        valid = np.zeros(n, np.uint8)
        flags = np.zeros(n, np.uint32)
        self._check(self.lib.mps_is_valid_batch(self.h, _fp(st), n, _fp(valid, C.c_uint8), _fp(flags, C.c_uint32)))
        return valid.astype(bool), flags

    # ---- distance helpers (host, double) ----------------------------------
    def distance(self, a, b, weights=None, mask=None) -> float:
        B = self.B
        out = C.c_double(0)
        full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
        w = _f32(weights, (B,))
        rc = self.lib.mps_distance(C.byref(self.scene), _fp(_f32(a, (3 * B,))), _fp(_f32(b, (3 * B,))), _fp(w),
                                   full if mask is None else int(mask), C.byref(out))
        self._check(rc)
        return out.value


class NearestNeighbors:
    """Device-resident tree with exact weighted nearest-neighbour queries.

    Mirrors the subset of ompl::NearestNeighbors<MotionPtr> the reference uses
    (pushing/algorithm/RRT.cpp:46-51, 202, 239, 757; algorithm/Interfaces.cpp:97-100):
    add, nearest, clear, size — plus parent / control / slice bookkeeping of Motion
    (ompl/planning/Essentials.cpp:20-93).  Ties resolve to the lowest index.
    """

    def __init__(self, ctx: Context, capacity: int = 0):
        self.ctx = ctx
        self.lib = ctx.lib
        if capacity:
            ctx._check(self.lib.mps_tree_reserve(ctx.h, capacity))

    def clear(self):
        self.ctx._check(self.lib.mps_tree_clear(self.ctx.h))

    def size(self) -> int:
        n = C.c_int64(0)
        self.ctx._check(self.lib.mps_tree_size(self.ctx.h, C.byref(n)))
        return n.value

    def add(self, state, parent: int = -1, control=None, slice_id: int = -1) -> int:
        idx = C.c_int64(0)
        s = _f32(state, (3 * self.ctx.B,))
        c = _f32(control, (5,))
        self.ctx._check(self.lib.mps_tree_add(self.ctx.h, _fp(s), parent, _fp(c), slice_id, C.byref(idx)))
        return idx.value

    def add_batch(self, states, parents=None, controls=None, slice_ids=None) -> int:
        st = _f32(states).reshape(-1, 3 * self.ctx.B)
        n = st.shape[0]
        pa = None if parents is None else np.ascontiguousarray(parents, np.int32)
        co = None if controls is None else _f32(controls).reshape(n, 5)
        sl = None if slice_ids is None else np.ascontiguousarray(slice_ids, np.int32)
        first = C.c_int64(0)
        self.ctx._check(self.lib.mps_tree_add_batch(self.ctx.h, _fp(st), _fp(pa, C.c_int32), _fp(co),
                                                    _fp(sl, C.c_int32), n, C.byref(first)))
        return first.value

    def nearest(self, query, weights=None, mask=None, slice_filter: int = -1):
        B = self.ctx.B
        full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
        idx = C.c_int64(-1)
        d = C.c_double(0)
        q = _f32(query, (3 * B,))
        w = _f32(weights, (B,))
        self.ctx._check(self.lib.mps_tree_nearest(self.ctx.h, _fp(q), _fp(w), full if mask is None else int(mask),
                                                  slice_filter, C.byref(idx), C.byref(d)))
        return idx.value, d.value

    def nearest_k(self, query, k: int, weights=None, mask=None, slice_filter: int = -1):
        """ompl nearestK: (indices [n], distances [n]) of the n <= k nearest nodes, ascending (double distance, index)"""
        B = self.ctx.B
        q = _f32(query, (3 * B,))
        w = _f32(weights, (B,))
        full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
        idx = np.zeros(k, np.int64)
        d = np.zeros(k, np.float64)
        n = C.c_int32(0)
        self.ctx._check(self.lib.mps_tree_nearest_k(self.ctx.h, _fp(q), _fp(w), full if mask is None else int(mask),
                                                    int(slice_filter), int(k), _fp(idx, C.c_int64), _fp(d, C.c_double),
                                                    C.byref(n)))
        return idx[: n.value], d[: n.value]

    def nearest_timed(self, queries, weights=None, mask=None, slice_filter: int = -1):
        """n single queries through the C entry point in a C loop -> (indices, distances, seconds)"""
        B = self.ctx.B
        qs = _f32(queries).reshape(-1, 3 * B)
        n = qs.shape[0]
        w = _f32(weights, (B,))
        full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
        idx = np.zeros(n, np.int64)
        d = np.zeros(n, np.float64)
        sec = C.c_double(0)
        self.ctx._check(self.lib.mps_tree_nearest_timed(self.ctx.h, _fp(qs), n, _fp(w), full if mask is None else int(mask),
                                                        int(slice_filter), _fp(idx, C.c_int64), _fp(d, C.c_double), C.byref(sec)))
        return idx, d, sec.value

    def nearest_k_fallbacks(self) -> int:
        n = C.c_int64(0)
        self.ctx._check(self.lib.mps_tree_nearest_k_fallbacks(self.ctx.h, C.byref(n)))
        return n.value

    def nearest_two_level(self, reps: "NearestNeighbors", query, rep_mask: int, member_mask: int, weights=None):
        """chained two-level search: ((rep_index, rep_distance), (index, distance)); reps = the tree of slice
        representatives (node index = slice id) in a second context on the same device"""
        B = self.ctx.B
        q = _f32(query, (3 * B,))
        w = _f32(weights, (B,))
        ri, rd, i, d = C.c_int64(-1), C.c_double(0), C.c_int64(-1), C.c_double(0)
        self.ctx._check(self.lib.mps_tree_nearest_two_level(self.ctx.h, reps.ctx.h, _fp(q), _fp(w), int(rep_mask), int(member_mask),
                                                            C.byref(ri), C.byref(rd), C.byref(i), C.byref(d)))
        return (ri.value, rd.value), (i.value, d.value)

    def nearest_batch(self, queries, weights=None, masks=None, slice_filters=None):
        B = self.ctx.B
        qs = _f32(queries).reshape(-1, 3 * B)
        Q = qs.shape[0]
        w = _f32(weights, (B,))
        m = None if masks is None else np.ascontiguousarray(masks, np.uint32)
        f = None if slice_filters is None else np.ascontiguousarray(slice_filters, np.int32)
        idx = np.zeros(Q, np.int64)
        d = np.zeros(Q, np.float64)
        self.ctx._check(self.lib.mps_tree_nearest_batch(self.ctx.h, _fp(qs), Q, _fp(w), _fp(m, C.c_uint32),
                                                        _fp(f, C.c_int32), _fp(idx, C.c_int64), _fp(d, C.c_double)))
        return idx, d

    def nearest_upload(self, queries, weights=None, masks=None, slice_filters=None):
        B = self.ctx.B
        qs = _f32(queries).reshape(-1, 3 * B)
        self._Q = qs.shape[0]
        w = _f32(weights, (B,))
        m = None if masks is None else np.ascontiguousarray(masks, np.uint32)
        f = None if slice_filters is None else np.ascontiguousarray(slice_filters, np.int32)
        self.ctx._check(self.lib.mps_tree_nearest_upload(self.ctx.h, _fp(qs), self._Q, _fp(w), _fp(m, C.c_uint32),
                                                         _fp(f, C.c_int32)))

    def nearest_launch(self):
        self.ctx._check(self.lib.mps_tree_nearest_launch(self.ctx.h))

    def nearest_fetch(self):
        idx = np.zeros(self._Q, np.int64)
        d = np.zeros(self._Q, np.float64)
        self.ctx._check(self.lib.mps_tree_nearest_fetch(self.ctx.h, _fp(idx, C.c_int64), _fp(d, C.c_double)))
        return idx, d

    def get(self, index: int):
        B = self.ctx.B
        s = np.zeros(3 * B, np.float32)
        c = np.zeros(5, np.float32)
        p = C.c_int32(0)
        sl = C.c_int32(0)
        self.ctx._check(self.lib.mps_tree_get(self.ctx.h, index, _fp(s), C.byref(p), _fp(c), C.byref(sl)))
        return s, p.value, c, sl.value

    def backtrack(self, index: int, max_len: int = 1 << 20) -> np.ndarray:
        buf = np.zeros(min(max_len, max(self.size(), 1)), np.int64)
        n = C.c_int64(0)
        self.ctx._check(self.lib.mps_tree_backtrack(self.ctx.h, index, _fp(buf, C.c_int64), len(buf), C.byref(n)))
        return buf[: n.value].copy()

    def fill_random(self, n: int, seed: int, n_slices: int = 0):
        self.ctx._check(self.lib.mps_tree_fill_random(self.ctx.h, n, seed, n_slices))

    def read_states(self, first: int, n: int) -> np.ndarray:
        out = np.zeros((n, 3 * self.ctx.B), np.float32)
        self.ctx._check(self.lib.mps_tree_read_states(self.ctx.h, first, n, _fp(out)))
        return out


# struct mps_extend_best as a numpy record (same layout as abi.ExtendBest)
_BEST_DTYPE = np.dtype([("k", np.int32), ("n_ok", np.int32), ("goal_step", np.int32), ("tree_index", np.int32),
                        ("distance", np.float64)])
assert _BEST_DTYPE.itemsize == C.sizeof(abi.ExtendBest)


class Forest:
    """Many independent planner instances in one context (mps_forest_*): one tree per instance in a shared
    dimension-major SoA, batches of M instances per launch (BASELINE config 5)."""

    def __init__(self, ctx: Context, n_instances: int, cap_per_instance: int):
        self.ctx = ctx
        self.lib = ctx.lib
        self.n = int(n_instances)
        ctx._check(self.lib.mps_forest_reserve(ctx.h, self.n, int(cap_per_instance)))

    def reset(self, roots):
        r = _f32(roots).reshape(self.n, 3 * self.ctx.B)
        self.ctx._check(self.lib.mps_forest_reset(self.ctx.h, _fp(r)))

    def sizes(self) -> np.ndarray:
        out = np.zeros(self.n, np.int64)
        self.ctx._check(self.lib.mps_forest_sizes(self.ctx.h, _fp(out, C.c_int64)))
        return out

    def nearest(self, inst_ids, queries, weights=None, masks=None):
        B = self.ctx.B
        ids = np.ascontiguousarray(inst_ids, np.int32)
        M = len(ids)
        qs = _f32(queries).reshape(M, 3 * B)
        w = _f32(weights, (B,))
        m = None if masks is None else np.ascontiguousarray(masks, np.uint32).reshape(M)
        idx = np.zeros(M, np.int64)
        d = np.zeros(M, np.float64)
        self.ctx._check(self.lib.mps_forest_nearest(self.ctx.h, M, _fp(ids, C.c_int32), _fp(qs), _fp(w),
                                                    _fp(m, C.c_uint32), _fp(idx, C.c_int64), _fp(d, C.c_double)))
        return idx, d

    def extend(self, inst_ids, parents, controls, n_ctrl=None, dests=None, masks=None, weights=None,
               compare_f32=False, goals=None, append=True):
        B = self.ctx.B
        ids = np.ascontiguousarray(inst_ids, np.int32)
        M = len(ids)
        par = np.ascontiguousarray(parents, np.int32).reshape(M)
        c = _f32(controls)
        if c.ndim == 3:
            c = c[:, :, None, :]
        assert c.shape[0] == M and c.shape[3] == 5
        K, L = c.shape[1], c.shape[2]
        nc = None if n_ctrl is None else np.ascontiguousarray(n_ctrl, np.int32).reshape(M, K)
        de = None if dests is None else _f32(dests).reshape(M, 3 * B)
        ma = None if masks is None else np.ascontiguousarray(masks, np.uint32).reshape(M)
        w = _f32(weights, (B,))
        garr, ng = None, 0
        if goals:
            ng = len(goals)
            garr = (abi.Goal * ng)()
            for i, g in enumerate(goals):
                garr[i].body, garr[i].x, garr[i].y, garr[i].tol = int(g[0]), float(g[1]), float(g[2]), float(g[3])
        fin = abi.ForestExtendIn()
        fin.M, fin.K, fin.L = M, K, L
        fin.inst_ids = _fp(ids, C.c_int32)
        fin.parents = _fp(par, C.c_int32)
        fin.controls = _fp(c)
        fin.n_ctrl = _fp(nc, C.c_int32)
        fin.dests = _fp(de)
        fin.masks = _fp(ma, C.c_uint32)
        fin.weights = _fp(w)
        fin.compare_f32 = 1 if compare_f32 else 0
        fin.n_goals = ng
        fin.goals = garr if garr is not None else None
        fin.append_to_tree = 1 if append else 0
        best = (abi.ExtendBest * M)()
        bs = np.zeros((M, L, 3 * B), np.float32)
        bc = np.zeros((M, L, 5), np.float32)
        bf = np.zeros((M, L), np.uint32)
        self.ctx._check(self.lib.mps_forest_extend(self.ctx.h, C.byref(fin), best, _fp(bs), _fp(bc), _fp(bf, C.c_uint32)))
        rec = np.frombuffer(best, dtype=_BEST_DTYPE, count=M)  # one structured view instead of M attribute reads
        return dict(k=rec["k"].copy(), n_ok=rec["n_ok"].copy(), goal_step=rec["goal_step"].copy(),
                    tree_index=rec["tree_index"].copy(), distance=rec["distance"].copy(), states=bs, controls=bc,
                    flags=bf)

    def naive_step(self, inst_ids, iteration: int, seed: int, goals, K: int = 10, C_: int = 8, goal_bias: float = 0.2,
                   target_bias: float = 0.25, robot_bias: float = 0.0, instance_offset: int = 0, weights=None,
                   velocity_limit: float = 0.6, omega_limit: float = 1.5, duration_limits=(0.1, 1.0)):
        """One NaiveRRT iteration of the given instances entirely on the device (mps_forest_naive_step).
        Returns dict(k, n_ok, goal_step, tree_index, distance, flags), one entry per instance."""
        ids = np.ascontiguousarray(inst_ids, np.int32)
        M = len(ids)
        garr = (abi.Goal * len(goals))()
        for i, g in enumerate(goals):
            garr[i].body, garr[i].x, garr[i].y, garr[i].tol = int(g[0]), float(g[1]), float(g[2]), float(g[3])
        w = None if weights is None else _f32(weights, (self.ctx.B,))
        fin = abi.ForestNaiveIn()
        fin.M, fin.iteration = M, int(iteration)
        fin.inst_ids = _fp(ids, C.c_int32)
        fin.instance_offset = int(instance_offset)
        fin.seed = int(seed)
        fin.K, fin.C = int(K), int(C_)
        fin.goal_bias, fin.target_bias, fin.robot_bias = float(goal_bias), float(target_bias), float(robot_bias)
        fin.n_goals = len(goals)
        fin.goals = garr
        fin.weights = _fp(w)
        fin.velocity_limit, fin.omega_limit = float(velocity_limit), float(omega_limit)
        fin.duration_min, fin.duration_max = float(duration_limits[0]), float(duration_limits[1])
        best = (abi.ExtendBest * M)()
        fl = np.zeros(M, np.uint32)
        self.ctx._check(self.lib.mps_forest_naive_step(self.ctx.h, C.byref(fin), best, _fp(fl, C.c_uint32)))
        rec = np.frombuffer(best, dtype=_BEST_DTYPE, count=M)
        return dict(k=rec["k"].copy(), n_ok=rec["n_ok"].copy(), goal_step=rec["goal_step"].copy(),
                    tree_index=rec["tree_index"].copy(), distance=rec["distance"].copy(), flags=fl)

    def read(self, inst: int, first: int, n: int):
        B = self.ctx.B
        st = np.zeros((n, 3 * B), np.float32)
        pa = np.zeros(n, np.int32)
        co = np.zeros((n, 5), np.float32)
        self.ctx._check(self.lib.mps_forest_read(self.ctx.h, int(inst), int(first), int(n), _fp(st), _fp(pa, C.c_int32),
                                                 _fp(co)))
        return st, pa, co
