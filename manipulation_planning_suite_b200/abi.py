"""ctypes mirror of include/mps_b200.h (the C ABI of the B200 extend step).

Only type definitions and the loader live here; no arithmetic.  The loader fails
loudly when the CUDA library has not been built — there is no CPU fallback
(BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import os

MPS_ABI_VERSION = 2
MPS_MAX_BODIES = 32
MPS_MAX_STATICS = 16
MPS_MAX_GOALS = 8
MPS_CONTROL_DIM = 5
MPS_MAX_MANIFOLDS = 64
MPS_MAX_CHAIN = 64

MPS_OK = 0
MPS_ERR_ARG = -1
MPS_ERR_NO_DEVICE = -2
MPS_ERR_CUDA = -3
MPS_ERR_CAPACITY = -4
MPS_ERR_STATE = -5

MPS_SHAPE_BOX = 0
MPS_SHAPE_DISC = 1

MPS_F_RAN = 1 << 0
MPS_F_OK = 1 << 1
MPS_F_GOAL = 1 << 2
MPS_F_IN_BALL = 1 << 3
MPS_F_NOT_AT_REST = 1 << 4
MPS_F_REST_CONTACT = 1 << 5
MPS_F_BOUNDS = 1 << 6
MPS_F_INFEASIBLE = 1 << 7
MPS_F_BAD_CONTACT = 1 << 8
MPS_F_OVERFLOW = 1 << 9
MPS_F_ACTIVE_CONTACT = 1 << 10
MPS_F_JAMMED = 1 << 11

_f32 = C.c_float
_i32 = C.c_int32
_u32 = C.c_uint32


class Scene(C.Structure):
    """struct mps_scene"""

    _fields_ = [
        ("n_bodies", _i32),
        ("robot_id", _i32),
        ("n_statics", _i32),
        ("vel_iters", _i32),
        ("shape", _i32 * MPS_MAX_BODIES),
        ("hx", _f32 * MPS_MAX_BODIES),
        ("hy", _f32 * MPS_MAX_BODIES),
        ("mass", _f32 * MPS_MAX_BODIES),
        ("inertia", _f32 * MPS_MAX_BODIES),
        ("mu_ground", _f32 * MPS_MAX_BODIES),
        ("fr_radius", _f32 * MPS_MAX_BODIES),
        ("mu_contact", _f32 * MPS_MAX_BODIES),
        ("s_shape", _i32 * MPS_MAX_STATICS),
        ("s_x", _f32 * MPS_MAX_STATICS),
        ("s_y", _f32 * MPS_MAX_STATICS),
        ("s_theta", _f32 * MPS_MAX_STATICS),
        ("s_hx", _f32 * MPS_MAX_STATICS),
        ("s_hy", _f32 * MPS_MAX_STATICS),
        ("s_mu_contact", _f32 * MPS_MAX_STATICS),
        ("x_min", _f32),
        ("x_max", _f32),
        ("y_min", _f32),
        ("y_max", _f32),
        ("pair_forbidden", _u32 * MPS_MAX_BODIES),
        ("static_forbidden", _u32),
        ("accel_limits", _f32 * 3),
        ("_pad0", _f32),
        ("position_weight", C.c_double),
        ("orientation_weight", C.c_double),
        ("dt", _f32),
        ("t_max", _f32),
        ("gravity", _f32),
        ("slop", _f32),
        ("baumgarte", _f32),
        ("max_bias_vel", _f32),
        ("pen_max", _f32),
        ("v_rest", _f32),
        ("w_rest", _f32),
        ("max_manifolds", _i32),
        ("strict_active_contacts", _i32),
        ("jam_steps", _i32),
    ]


class Goal(C.Structure):
    """struct mps_goal"""

    _fields_ = [("body", _i32), ("x", _f32), ("y", _f32), ("tol", _f32)]


class ExtendIn(C.Structure):
    """struct mps_extend_in"""

    _fields_ = [
        ("start", C.POINTER(_f32)),
        ("controls", C.POINTER(_f32)),
        ("n_ctrl", C.POINTER(_i32)),
        ("K", _i32),
        ("L", _i32),
        ("dest", C.POINTER(_f32)),
        ("weights", C.POINTER(_f32)),
        ("mask", _u32),
        ("compare_f32", _i32),
        ("goals", C.POINTER(Goal)),
        ("n_goals", _i32),
        ("k_offset", _i32),
        ("ball_center", C.POINTER(_f32)),
        ("ball_radius", _f32),
        ("append_to_tree", _i32),
        ("parent", _i32),
        ("stop_at_goal", _i32),
        ("starts", C.POINTER(_f32)),
    ]


class ExtendBest(C.Structure):
    """struct mps_extend_best"""

    _fields_ = [
        ("k", _i32),
        ("n_ok", _i32),
        ("goal_step", _i32),
        ("tree_index", _i32),
        ("distance", C.c_double),
    ]


class HybridGen(C.Structure):
    """struct mps_hybrid_gen"""

    _fields_ = [
        ("seed", C.c_uint64),
        ("iteration", _i32),
        ("k_offset", _i32),
        ("fixed_object", _i32),
        ("action_randomness", _f32),
        ("target_bias", _f32),
        ("robot_bias", _f32),
        ("n_targets", _i32),
        ("targets", _i32 * MPS_MAX_GOALS),
        ("velocity_limits", _f32 * 3),
        ("duration_min", _f32),
        ("duration_max", _f32),
        ("optimal_push_distance", _f32),
        ("push_distance_tolerance", _f32),
        ("push_angle_tolerance", _f32),
        ("object_size", _f32 * MPS_MAX_BODIES),
    ]


class ForestNaiveIn(C.Structure):
    """struct mps_forest_naive_in"""

    _fields_ = [
        ("M", _i32), ("iteration", _i32),
        ("inst_ids", C.POINTER(_i32)),
        ("instance_offset", C.c_int64),
        ("seed", C.c_uint64),
        ("K", _i32), ("C", _i32),
        ("goal_bias", _f32), ("target_bias", _f32), ("robot_bias", _f32),
        ("n_goals", _i32),
        ("goals", C.POINTER(Goal)),
        ("weights", C.POINTER(_f32)),
        ("velocity_limit", _f32), ("omega_limit", _f32), ("duration_min", _f32), ("duration_max", _f32),
    ]


class ForestExtendIn(C.Structure):
    """struct mps_forest_extend_in"""

    _fields_ = [
        ("M", _i32), ("K", _i32), ("L", _i32),
        ("inst_ids", C.POINTER(_i32)),
        ("parents", C.POINTER(_i32)),
        ("controls", C.POINTER(_f32)),
        ("n_ctrl", C.POINTER(_i32)),
        ("dests", C.POINTER(_f32)),
        ("masks", C.POINTER(_u32)),
        ("weights", C.POINTER(_f32)),
        ("compare_f32", _i32),
        ("n_goals", _i32),
        ("goals", C.POINTER(Goal)),
        ("append_to_tree", _i32),
        ("_pad", _i32),
    ]


class ExtendOut(C.Structure):
    """struct mps_extend_out"""

    _fields_ = [
        ("best", C.POINTER(ExtendBest)),
        ("best_states", C.POINTER(_f32)),
        ("best_controls", C.POINTER(_f32)),
        ("best_flags", C.POINTER(_u32)),
        ("all_states", C.POINTER(_f32)),
        ("all_rest", C.POINTER(_f32)),
        ("all_flags", C.POINTER(_u32)),
        ("all_dist", C.POINTER(C.c_double)),
        ("all_steps", C.POINTER(_i32)),
    ]


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libmps_b200.so")

_lib = None

# name -> (restype, argtypes); every symbol include/mps_b200.h declares
_P = C.POINTER
_ctx = C.c_void_p
SYMBOLS = {
    "mps_abi_version": (C.c_int, []),
    "mps_last_error": (C.c_char_p, [_ctx]),
    "mps_device_count": (C.c_int, [_P(C.c_int)]),
    "mps_scene_defaults": (None, [_P(Scene)]),
    "mps_ctx_create": (C.c_int, [_P(Scene), C.c_int, _P(_ctx)]),
    "mps_ctx_create_on_stream": (C.c_int, [_P(Scene), C.c_int, C.c_void_p, _P(_ctx)]),
    "mps_ctx_destroy": (None, [_ctx]),
    "mps_ctx_synchronize": (C.c_int, [_ctx]),
    "mps_ctx_scene": (C.c_int, [_ctx, _P(Scene)]),
    "mps_extend": (C.c_int, [_ctx, _P(ExtendIn), _P(ExtendOut)]),
    "mps_propagate": (C.c_int, [_ctx, _P(_f32), _P(_f32), _P(_f32), _P(_i32), _P(_u32)]),
    "mps_extend_upload": (C.c_int, [_ctx, _P(ExtendIn)]),
    "mps_extend_upload_generated": (C.c_int, [_ctx, _P(ExtendIn), _P(HybridGen)]),
    "mps_extend_generated_controls": (C.c_int, [_ctx, _P(_f32), _P(_i32), _P(_i32)]),
    "mps_sample_controls": (C.c_int, [_ctx, C.c_uint64, _i32, C.c_int64, _i32, _f32, _f32, _f32, _f32, _P(_f32)]),
    "mps_extend_launch": (C.c_int, [_ctx]),
    "mps_extend_fetch": (C.c_int, [_ctx, _P(ExtendOut)]),
    "mps_extend_counters": (C.c_int, [_ctx, _P(C.c_uint64)]),
    "mps_extend_chain_cycles": (C.c_int, [_ctx, _P(C.c_int64)]),
    "mps_is_valid_batch": (C.c_int, [_ctx, _P(_f32), _i32, _P(C.c_uint8), _P(_u32)]),
    "mps_distance": (C.c_int, [_P(Scene), _P(_f32), _P(_f32), _P(_f32), _u32, _P(C.c_double)]),
    "mps_goal_distance": (C.c_int, [_P(Scene), _P(_f32), _P(Goal), _i32, _P(C.c_double)]),
    "mps_tree_reserve": (C.c_int, [_ctx, C.c_int64]),
    "mps_tree_clear": (C.c_int, [_ctx]),
    "mps_tree_size": (C.c_int, [_ctx, _P(C.c_int64)]),
    "mps_tree_add": (C.c_int, [_ctx, _P(_f32), _i32, _P(_f32), _i32, _P(C.c_int64)]),
    "mps_tree_add_batch": (C.c_int, [_ctx, _P(_f32), _P(_i32), _P(_f32), _P(_i32), C.c_int64, _P(C.c_int64)]),
    "mps_tree_nearest": (C.c_int, [_ctx, _P(_f32), _P(_f32), _u32, _i32, _P(C.c_int64), _P(C.c_double)]),
    "mps_tree_nearest_k": (C.c_int, [_ctx, _P(_f32), _P(_f32), _u32, _i32, _i32, _P(C.c_int64), _P(C.c_double), _P(_i32)]),
    "mps_tree_nearest_timed": (C.c_int, [_ctx, _P(_f32), _i32, _P(_f32), _u32, _i32, _P(C.c_int64), _P(C.c_double), _P(C.c_double)]),
    "mps_tree_nearest_k_fallbacks": (C.c_int, [_ctx, _P(C.c_int64)]),
    "mps_tree_nearest_two_level": (C.c_int, [_ctx, _ctx, _P(_f32), _P(_f32), _u32, _u32, _P(C.c_int64), _P(C.c_double),
                                             _P(C.c_int64), _P(C.c_double)]),
    "mps_tree_nearest_batch": (
        C.c_int,
        [_ctx, _P(_f32), _i32, _P(_f32), _P(_u32), _P(_i32), _P(C.c_int64), _P(C.c_double)],
    ),
    "mps_tree_nearest_upload": (C.c_int, [_ctx, _P(_f32), _i32, _P(_f32), _P(_u32), _P(_i32)]),
    "mps_tree_nearest_launch": (C.c_int, [_ctx]),
    "mps_tree_nearest_fetch": (C.c_int, [_ctx, _P(C.c_int64), _P(C.c_double)]),
    "mps_tree_get": (C.c_int, [_ctx, C.c_int64, _P(_f32), _P(_i32), _P(_f32), _P(_i32)]),
    "mps_tree_backtrack": (C.c_int, [_ctx, C.c_int64, _P(C.c_int64), C.c_int64, _P(C.c_int64)]),
    "mps_tree_fill_random": (C.c_int, [_ctx, C.c_int64, C.c_uint64, _i32]),
    "mps_tree_read_states": (C.c_int, [_ctx, C.c_int64, C.c_int64, _P(_f32)]),
    "mps_time_launches": (C.c_int, [_ctx, _i32, _i32, _P(_f32)]),
    "mps_forest_reserve": (C.c_int, [_ctx, _i32, _i32]),
    "mps_forest_reset": (C.c_int, [_ctx, _P(_f32)]),
    "mps_forest_sizes": (C.c_int, [_ctx, _P(C.c_int64)]),
    "mps_forest_nearest": (C.c_int, [_ctx, _i32, _P(_i32), _P(_f32), _P(_f32), _P(_u32), _P(C.c_int64), _P(C.c_double)]),
    "mps_forest_extend": (C.c_int, [_ctx, _P(ForestExtendIn), _P(ExtendBest), _P(_f32), _P(_f32), _P(_u32)]),
    "mps_forest_naive_step": (C.c_int, [_ctx, _P(ForestNaiveIn), _P(ExtendBest), _P(_u32)]),
    "mps_forest_read": (C.c_int, [_ctx, _i32, C.c_int64, C.c_int64, _P(_f32), _P(_i32), _P(_f32)]),
    "mps_extend_record_size": (C.c_int, [_ctx, _P(C.c_int64)]),
    "mps_extend_record_to": (C.c_int, [_ctx, C.c_void_p]),
    "mps_extend_merge_records": (C.c_int, [_ctx, C.c_void_p, _i32, _i32, _i32, _i32, _i32, _P(ExtendBest)]),
    "mps_extend_head": (C.c_int, [_ctx, _P(ExtendBest)]),
    "mps_ctx_set_kernel_timing": (C.c_int, [_ctx, _i32]),
    "mps_extend_kernel_times": (C.c_int, [_ctx, _P(_f32), _i32, _P(_i32)]),
    "mps_measure_fp32_peak": (C.c_int, [_ctx, _P(_f32)]),
    "mps_ctx_stream": (C.c_int, [_ctx, _P(C.c_void_p)]),
}


class LibraryMissing(RuntimeError):
    pass


def load_library(path: str | None = None):
    """dlopen libmps_b200.so and type every entry point.  No fallback of any kind."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MPS_B200_LIB") or LIB_PATH  # MPS_B200_LIB: tuning builds (profiles/tools)
    if not os.path.exists(p):
        raise LibraryMissing(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The extend step has no CPU fallback."
        )
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mps_abi_version() != MPS_ABI_VERSION:
        raise RuntimeError("libmps_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
