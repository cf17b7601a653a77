"""ctypes binding of tests/emu/libmps_emu.so — TEST INFRASTRUCTURE (see tests/emu/cuda_emu.h): the repo's CUDA
kernel sources compiled for the host through a SIMT emulator, so that the kernels' source can be checked
against the oracle on a machine without a GPU.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import oracle_binding as ob
from manipulation_planning_suite_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-s", "-C", d])
        L = C.CDLL(os.path.join(d, "libmps_emu.so"))
        L.emu_extend.argtypes = [C.POINTER(abi.Scene), C.POINTER(abi.ExtendIn), C.POINTER(abi.ExtendOut), C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.emu_is_valid_batch.argtypes = [C.POINTER(abi.Scene), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint8),
                                         C.POINTER(C.c_uint32)]
        L.emu_nearest_k.argtypes = [C.POINTER(abi.Scene), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_longlong,
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32, C.c_int32, C.c_int, C.c_int,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.emu_nearest_next.restype = C.c_longlong
        L.emu_nearest_next.argtypes = [C.POINTER(abi.Scene), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_longlong,
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32, C.c_int32, C.c_double,
                                       C.c_longlong, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def extend(sc, start, controls, group=32, minb=3, mcap_small=0, ctas=2, counters=None, **kw):
    """rollout_kernel<., minb, group> + select_kernel through the emulator; same result dict as ob.extend"""
    cnt = (C.c_uint64 * 8)()

    def call(sc_, ein, eout):
        return lib().emu_extend(C.byref(sc_), C.byref(ein), C.byref(eout), group, minb, mcap_small, ctas, cnt)

    r = ob.extend(sc, start, controls, _call=call, **kw)
    r["counters"] = dict(steps=cnt[0], candidates=cnt[1], touching=cnt[2], colour_passes=cnt[3])
    return r


def is_valid_batch(sc, states):
    st = ob.f32(states).reshape(-1, 3 * sc.n_bodies)
    n = st.shape[0]
    valid = np.zeros(n, np.uint8)
    flags = np.zeros(n, np.uint32)
    lib().emu_is_valid_batch(C.byref(sc), ob.fp(st), n, valid.ctypes.data_as(C.POINTER(C.c_uint8)),
                             flags.ctypes.data_as(C.POINTER(C.c_uint32)))
    return valid.astype(bool), flags


def nearest_k(sc, states, query, k, weights=None, mask=None, slice_ids=None, slice_filter=-1, grid_x=3):
    """nn_topk_kernel through the emulator -> (indices, distances, status)"""
    B = sc.n_bodies
    st = ob.f32(states).reshape(-1, 3 * B)
    q = ob.f32(query).reshape(3 * B)
    w = None if weights is None else ob.f32(weights)
    sl = None if slice_ids is None else np.ascontiguousarray(slice_ids, np.int32)
    full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
    idx = np.zeros(32, np.int64)
    d = np.zeros(32, np.float64)
    status = C.c_int(0)
    n = lib().emu_nearest_k(C.byref(sc), ob.fp(st), None if sl is None else sl.ctypes.data_as(C.POINTER(C.c_int32)), st.shape[0],
                            ob.fp(q), None if w is None else ob.fp(w), full if mask is None else int(mask), int(slice_filter),
                            int(k), int(grid_x), idx.ctypes.data_as(C.POINTER(C.c_int64)),
                            d.ctypes.data_as(C.POINTER(C.c_double)), C.byref(status))
    return idx[:n], d[:n], status.value


def nearest_next(sc, states, query, after_d, after_i, weights=None, mask=None, slice_ids=None, slice_filter=-1, grid_x=2):
    B = sc.n_bodies
    st = ob.f32(states).reshape(-1, 3 * B)
    q = ob.f32(query).reshape(3 * B)
    w = None if weights is None else ob.f32(weights)
    sl = None if slice_ids is None else np.ascontiguousarray(slice_ids, np.int32)
    full = (1 << B) - 1 if B < 32 else 0xFFFFFFFF
    d = C.c_double(0)
    i = lib().emu_nearest_next(C.byref(sc), ob.fp(st), None if sl is None else sl.ctypes.data_as(C.POINTER(C.c_int32)),
                               st.shape[0], ob.fp(q), None if w is None else ob.fp(w), full if mask is None else int(mask),
                               int(slice_filter), float(after_d), int(after_i), int(grid_x), C.byref(d))
    return int(i), d.value
