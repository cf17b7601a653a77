"""ctypes binding of libmps_host.so: the C++ planner classes (host/mps_host.hpp, mirroring the
reference's RearrangementRRT family and OraclePushPlanner) driven through mps_host_plan()."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import abi

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_PKG_DIR, "host", "libmps_host.so")

ALGORITHMS = {"Naive": 0, "OracleRRT": 1, "SliceOracleRRT": 2, "HybridActionRRT": 3}
# PlanningProblem::ShortcutType (include/mps/planner/pushing/OraclePushPlanner.h:48-54)
SHORTCUT_TYPES = {"NaiveShortcut": 0, "OracleShortcut": 1, "NoShortcut": 2, "LocalShortcut": 3, "LocalOracleShortcut": 4}


class HostParams(C.Structure):
    _fields_ = [
        ("algorithm", C.c_int32), ("time_out", C.c_float), ("goal_bias", C.c_float), ("target_bias", C.c_float),
        ("robot_bias", C.c_float), ("num_control_samples", C.c_int32), ("action_randomness", C.c_float),
        ("state_noise", C.c_float), ("do_slice_ball_projection", C.c_int32), ("device", C.c_int32),
        ("seed", C.c_uint64), ("max_iterations", C.c_int32), ("velocity_limits", C.c_float * 3),
        ("duration_limits", C.c_float * 2), ("shortcut_type", C.c_int32), ("max_shortcut_time", C.c_float),
        ("shortcut_batch_size", C.c_int32),
    ]


class HostResult(C.Structure):
    _fields_ = [
        ("solved", C.c_int32), ("reproducible", C.c_int32), ("num_iterations", C.c_int32),
        ("num_state_propagations", C.c_int32), ("num_samples", C.c_int32), ("num_nearest_neighbor_queries", C.c_int32),
        ("runtime_cpu", C.c_float), ("runtime_wall", C.c_float), ("tree_size", C.c_int32), ("num_slices", C.c_int32),
        ("path_len", C.c_int32), ("cost_before_shortcut", C.c_float), ("cost_after_shortcut", C.c_float),
        ("reproducible_after_shortcut", C.c_int32), ("path_len_before_shortcut", C.c_int32),
        ("shortcut_candidates", C.c_int32), ("shortcut_launches", C.c_int32), ("shortcut_propagations", C.c_int32),
        ("shortcut_accepted", C.c_int32),
    ]


_lib = None


def load_host_library():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise abi.LibraryMissing(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
        abi.load_library()  # libmps_b200.so first (rpath also resolves it)
        L = C.CDLL(HOST_LIB_PATH)
        L.mps_host_default_params.argtypes = [C.POINTER(HostParams)]
        L.mps_host_plan.argtypes = [C.POINTER(abi.Scene), C.POINTER(C.c_float), C.POINTER(abi.Goal), C.c_int32,
                                    C.POINTER(HostParams), C.POINTER(HostResult), C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.c_int32, C.c_char_p, C.c_int32]
        L.mps_host_shortcut.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.Goal), C.c_int32, C.POINTER(HostParams),
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.POINTER(HostResult),
                                        C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_char_p, C.c_int32]
        L.mps_host_ramp_steer.argtypes = [C.POINTER(C.c_float)] * 6 + [C.c_int32]
        L.mps_host_max_pushing_distance.restype = C.c_float
        L.mps_host_project_slice_ball.argtypes = [C.POINTER(abi.Scene), C.POINTER(C.c_float), C.c_int32, C.c_float,
                                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib = L
    return _lib


@dataclass
class PlanResult:
    solved: bool
    reproducible: bool
    stats: dict
    path_states: np.ndarray
    path_controls: np.ndarray


def default_params() -> HostParams:
    p = HostParams()
    load_host_library().mps_host_default_params(C.byref(p))
    return p


def _params(algorithm, shortcut_type, kw) -> HostParams:
    p = default_params()
    p.algorithm = ALGORITHMS[algorithm] if isinstance(algorithm, str) else int(algorithm)
    p.shortcut_type = SHORTCUT_TYPES[shortcut_type] if isinstance(shortcut_type, str) else int(shortcut_type)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown planner parameter {k}")
        setattr(p, k, v)
    return p


def _goals(goals):
    g = (abi.Goal * len(goals))()
    for i, (body, x, y, tol) in enumerate(goals):
        g[i].body, g[i].x, g[i].y, g[i].tol = int(body), float(x), float(y), float(tol)
    return g


def plan(scene: abi.Scene, start, goals, algorithm="Naive", max_path=4096, shortcut_type="NoShortcut",
         **kw) -> PlanResult:
    """OraclePushPlanner::setup + solve (pushing/OraclePushPlanner.cpp:107-243) on the device.  `shortcut_type`
    defaults to NoShortcut here (the reference's PlanningProblem defaults to LocalShortcut, :87)."""
    L = load_host_library()
    p = _params(algorithm, shortcut_type, kw)
    B = scene.n_bodies
    s = np.ascontiguousarray(start, np.float32).reshape(3 * B)
    g = _goals(goals)
    res = HostResult()
    ps = np.zeros((max_path, 3 * B), np.float32)
    pc = np.zeros((max_path, 5), np.float32)
    err = C.create_string_buffer(512)
    rc = L.mps_host_plan(C.byref(scene), s.ctypes.data_as(C.POINTER(C.c_float)), g, len(goals), C.byref(p), C.byref(res),
                         ps.ctypes.data_as(C.POINTER(C.c_float)), pc.ctypes.data_as(C.POINTER(C.c_float)), max_path,
                         err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    n = min(res.path_len, max_path)
    stats = {f: getattr(res, f) for f, _ in HostResult._fields_}
    return PlanResult(bool(res.solved), bool(res.reproducible), stats, ps[:n].copy(), pc[:n].copy())


def shortcut(scene: abi.Scene, goals, path_states, path_controls, shortcut_type="LocalShortcut", max_path=4096,
             **kw) -> PlanResult:
    """Shortcuts a solution path (the shortcut step of OraclePushPlanner::solve, :229-238); every round of
    candidates is propagated by batched launches."""
    L = load_host_library()
    p = _params("Naive", shortcut_type, kw)
    B = scene.n_bodies
    ins = np.ascontiguousarray(path_states, np.float32).reshape(-1, 3 * B)
    inc = np.ascontiguousarray(path_controls, np.float32).reshape(-1, 5)
    assert len(ins) == len(inc)
    g = _goals(goals)
    res = HostResult()
    ps = np.zeros((max_path, 3 * B), np.float32)
    pc = np.zeros((max_path, 5), np.float32)
    err = C.create_string_buffer(512)
    P = C.POINTER(C.c_float)
    rc = L.mps_host_shortcut(C.byref(scene), g, len(goals), C.byref(p), ins.ctypes.data_as(P), inc.ctypes.data_as(P),
                             len(ins), C.byref(res), ps.ctypes.data_as(P), pc.ctypes.data_as(P), max_path, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    n = min(res.path_len, max_path)
    stats = {f: getattr(res, f) for f, _ in HostResult._fields_}
    return PlanResult(bool(res.solved), bool(res.reproducible), stats, ps[:n].copy(), pc[:n].copy())


def ramp_steer(current, desired, vel_limits=(0.6, 0.6, 1.5), accel_limits=(2.0, 2.0, 4.0), duration_limits=(0.1, 1.0)):
    L = load_host_library()
    f = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
    v, a, d, c, e = f(vel_limits), f(accel_limits), f(duration_limits), f(current), f(desired)
    out = np.zeros((64, 5), np.float32)
    P = C.POINTER(C.c_float)
    n = L.mps_host_ramp_steer(v.ctypes.data_as(P), a.ctypes.data_as(P), d.ctypes.data_as(P), c.ctypes.data_as(P),
                              e.ctypes.data_as(P), out.ctypes.data_as(P), 64)
    return out[:n].copy()


def project_slice_ball(scene: abi.Scene, center, sample, radius: float, weights=None):
    """SliceBasedOracleRRT::projectSliceBall -> (inside, projected sample)"""
    B = int(scene.n_bodies)
    c = np.ascontiguousarray(center, np.float32).reshape(3 * B)
    s = np.ascontiguousarray(sample, np.float32).reshape(3 * B).copy()
    w = None if weights is None else np.ascontiguousarray(weights, np.float32)
    fp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    rc = load_host_library().mps_host_project_slice_ball(C.byref(scene), fp(w), int(scene.robot_id), float(radius), fp(c), fp(s))
    if rc < 0:
        raise ValueError("bad argument to mps_host_project_slice_ball")
    return bool(rc), s
