"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/mps_oracle.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from manipulation_planning_suite_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_P = C.POINTER
_f = C.c_float


class Ramp(C.Structure):
    _fields_ = [("v", _f * 3), ("a", _f * 3), ("t_acc", _f), ("t_plateau", _f), ("t_rest", _f)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.orc_normalize_orientation_d.argtypes = [_P(C.c_double)]
        L.orc_normalize_orientation_f.argtypes = [_P(_f)]
        L.orc_shortest_direction_so2.restype = _f
        L.orc_shortest_direction_so2.argtypes = [_f, _f]
        L.orc_ramp_set.argtypes = [_P(Ramp), _P(_f), _P(_f)]
        L.orc_ramp_velocity.argtypes = [_P(Ramp), _f, _P(_f)]
        L.orc_ramp_max_duration.restype = _f
        L.orc_ramp_max_duration.argtypes = [_P(Ramp)]
        L.orc_ramp_from_local_twist.argtypes = [_P(_f), _P(_f), _P(_f)]
        L.orc_distance.restype = C.c_double
        L.orc_distance.argtypes = [_P(abi.Scene), _P(_f), _P(_f), _P(_f), C.c_uint32]
        L.orc_goal_distance.restype = C.c_double
        L.orc_goal_distance.argtypes = [_P(_f), _P(abi.Goal), C.c_int]
        L.orc_goal_satisfied.argtypes = [_P(_f), _P(abi.Goal), C.c_int]
        L.orc_collision_allowed.argtypes = [_P(abi.Scene), C.c_int, C.c_int]
        L.orc_sincos.argtypes = [_f, _P(_f), _P(_f)]
        L.orc_is_valid.argtypes = [_P(abi.Scene), _P(_f), _P(C.c_uint32)]
        L.orc_propagate.argtypes = [_P(abi.Scene), _P(_f), _P(_f), _P(_f), _P(C.c_uint32), _P(C.c_int32)]
        L.orc_extend.argtypes = [_P(abi.Scene), _P(abi.ExtendIn), _P(abi.ExtendOut)]
        L.orc_extend_mt.argtypes = [_P(abi.Scene), _P(abi.ExtendIn), _P(abi.ExtendOut), C.c_int]
        L.orc_nn_linear.restype = C.c_int64
        L.orc_nn_linear.argtypes = [_P(abi.Scene), _P(_f), _P(C.c_int32), C.c_int64, _P(_f), _P(_f), C.c_uint32,
                                    C.c_int32, _P(C.c_double)]
        L.orc_nn_linear_batch_mt.argtypes = [_P(abi.Scene), _P(_f), _P(C.c_int32), C.c_int64, _P(_f), C.c_int32,
                                             _P(_f), _P(C.c_uint32), _P(C.c_int32), _P(C.c_int64), _P(C.c_double),
                                             C.c_int]
        L.orc_propagate_trace.argtypes = [_P(abi.Scene), _P(_f), _P(_f), _P(_f), C.c_int32, _P(C.c_int32)]
        L.orc_step_once.argtypes = [_P(abi.Scene), _P(_f), _P(_f), _P(_f), _P(C.c_int32)]
        L.orc_contacts.argtypes = [_P(abi.Scene), _P(_f), _P(C.c_int32), _P(_f), C.c_int32]
        L.orc_gnat_create.restype = C.c_void_p
        L.orc_gnat_create.argtypes = [_P(abi.Scene), _P(_f), C.c_uint32]
        L.orc_gnat_free.argtypes = [C.c_void_p]
        L.orc_gnat_add.argtypes = [C.c_void_p, _P(_f)]
        L.orc_gnat_size.argtypes = [C.c_void_p]
        L.orc_gnat_nearest.argtypes = [C.c_void_p, _P(_f), _P(C.c_double)]
        L.orc_gnat_distance_evaluations.restype = C.c_long
        L.orc_gnat_distance_evaluations.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class Gnat:
    """orc_gnat: CPU restatement of ompl::NearestNeighborsGNAT (oracle/mps_gnat.c)"""

    def __init__(self, sc, weights=None, mask=None):
        B = sc.n_bodies
        full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
        self.sc = sc
        w = f32(weights)
        self.h = lib().orc_gnat_create(C.byref(sc), fp(w), full if mask is None else int(mask))

    def add(self, state):
        s = f32(state)
        return lib().orc_gnat_add(self.h, fp(s))

    def add_many(self, states):
        st = f32(states).reshape(-1, 3 * self.sc.n_bodies)
        for i in range(len(st)):
            lib().orc_gnat_add(self.h, fp(st[i]))

    def nearest(self, q):
        d = C.c_double(0)
        qq = f32(q)
        i = lib().orc_gnat_nearest(self.h, fp(qq), C.byref(d))
        return int(i), d.value

    def distance_evaluations(self):
        return lib().orc_gnat_distance_evaluations(self.h)

    def __del__(self):
        try:
            lib().orc_gnat_free(self.h)
        except Exception:
            pass


def fp(a, t=_f):
    return None if a is None else a.ctypes.data_as(_P(t))


def f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def propagate(sc, state, control):
    B = sc.n_bodies
    s = f32(state).reshape(3 * B).copy()
    c = f32(control).reshape(5).copy()
    out = np.zeros(3 * B, np.float32)
    fl = C.c_uint32(0)
    st = C.c_int32(0)
    ok = lib().orc_propagate(C.byref(sc), fp(s), fp(c), fp(out), C.byref(fl), C.byref(st))
    return bool(ok), out, c, int(fl.value), int(st.value)


def propagate_trace(sc, state, control, max_steps=2000):
    B = sc.n_bodies
    s = f32(state).reshape(3 * B).copy()
    c = f32(control).reshape(5).copy()
    tr = np.zeros((max_steps, B, 6), np.float32)
    n = C.c_int32(0)
    ok = lib().orc_propagate_trace(C.byref(sc), fp(s), fp(c), fp(tr), max_steps, C.byref(n))
    return bool(ok), tr[: n.value], c


def is_valid(sc, state):
    fl = C.c_uint32(0)
    s = f32(state).copy()
    ok = lib().orc_is_valid(C.byref(sc), fp(s), C.byref(fl))
    return bool(ok), int(fl.value)


def distance(sc, a, b, weights=None, mask=None):
    B = sc.n_bodies
    full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
    a = f32(a)
    b = f32(b)
    w = f32(weights)
    return lib().orc_distance(C.byref(sc), fp(a), fp(b), fp(w), full if mask is None else int(mask))


def contacts(sc, state, max_m=64):
    s = f32(state).copy()
    ab = np.zeros((max_m, 2), np.int32)
    data = np.zeros((max_m, 5), np.float32)
    n = lib().orc_contacts(C.byref(sc), fp(s), fp(ab, C.c_int32), fp(data), max_m)
    return ab[:n], data[:n]


def extend(sc, start, controls, n_ctrl=None, dest=None, weights=None, mask=None, compare_f32=False, goals=None,
           k_offset=0, ball_center=None, ball_radius=0.0, threads=0, starts=None, stop_at_goal=False, _call=None):
    """Same inputs / outputs as manipulation_planning_suite_b200.api.Context.extend(dump=True).
    _call(sc, ein, eout) -> rc: another implementation taking the same structs (tests/emu_binding.py)."""
    B = sc.n_bodies
    controls = f32(controls)
    if controls.ndim == 2:
        controls = controls[:, None, :]
    K, L, _ = controls.shape
    start = None if start is None else f32(start).reshape(3 * B)
    starts_a = None if starts is None else f32(starts).reshape(K, 3 * B)
    dest = None if dest is None else f32(dest).reshape(3 * B)
    weights = None if weights is None else f32(weights).reshape(B)
    ball_center = None if ball_center is None else f32(ball_center).reshape(3 * B)
    n_ctrl_a = None if n_ctrl is None else np.ascontiguousarray(n_ctrl, np.int32)
    garr = None
    ng = 0
    if goals:
        ng = len(goals)
        garr = (abi.Goal * ng)()
        for i, g in enumerate(goals):
            garr[i].body, garr[i].x, garr[i].y, garr[i].tol = int(g[0]), float(g[1]), float(g[2]), float(g[3])
    ein = abi.ExtendIn()
    ein.start = fp(start)
    ein.controls = fp(controls)
    ein.n_ctrl = fp(n_ctrl_a, C.c_int32)
    ein.K, ein.L = K, L
    ein.dest = fp(dest)
    ein.weights = fp(weights)
    full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
    ein.mask = full if mask is None else int(mask)
    ein.compare_f32 = 1 if compare_f32 else 0
    ein.goals = garr if garr is not None else None
    ein.n_goals = ng
    ein.k_offset = k_offset
    ein.ball_center = fp(ball_center)
    ein.ball_radius = ball_radius
    ein.starts = fp(starts_a)
    ein.stop_at_goal = 1 if stop_at_goal else 0
    best = abi.ExtendBest()
    res = dict(
        best=best,
        best_states=np.zeros((L, 3 * B), np.float32),
        best_controls=np.zeros((L, 5), np.float32),
        best_flags=np.zeros(L, np.uint32),
        all_states=np.zeros((K, L, 3 * B), np.float32),
        all_rest=np.zeros((K, L), np.float32),
        all_flags=np.zeros((K, L), np.uint32),
        all_dist=np.zeros(K, np.float64),
        all_steps=np.zeros((K, L), np.int32),
    )
    eout = abi.ExtendOut()
    eout.best = C.pointer(best)
    eout.best_states = fp(res["best_states"])
    eout.best_controls = fp(res["best_controls"])
    eout.best_flags = fp(res["best_flags"], C.c_uint32)
    eout.all_states = fp(res["all_states"])
    eout.all_rest = fp(res["all_rest"])
    eout.all_flags = fp(res["all_flags"], C.c_uint32)
    eout.all_dist = fp(res["all_dist"], C.c_double)
    eout.all_steps = fp(res["all_steps"], C.c_int32)
    if _call is not None:
        rc = _call(sc, ein, eout)
    elif threads and threads > 1:
        rc = lib().orc_extend_mt(C.byref(sc), C.byref(ein), C.byref(eout), threads)
    else:
        rc = lib().orc_extend(C.byref(sc), C.byref(ein), C.byref(eout))
    assert rc == 0
    res["k"], res["n_ok"], res["goal_step"], res["distance"] = best.k, best.n_ok, best.goal_step, best.distance
    return res


def nn_linear(sc, states, query, weights=None, mask=None, slice_ids=None, slice_filter=-1):
    B = sc.n_bodies
    st = f32(states).reshape(-1, 3 * B)
    q = f32(query).reshape(3 * B)
    w = f32(weights)
    sl = None if slice_ids is None else np.ascontiguousarray(slice_ids, np.int32)
    full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
    d = C.c_double(0)
    i = lib().orc_nn_linear(C.byref(sc), fp(st), fp(sl, C.c_int32), st.shape[0], fp(q), fp(w),
                            full if mask is None else int(mask), slice_filter, C.byref(d))
    return int(i), d.value
