"""ctypes wrapper of the C oracle (oracle/dtsdf_oracle.c).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "dtsdf_oracle.c")
LIB = os.path.join(_HERE, "_build", "libdtsdf_oracle.so")

_lib = None


def build(force=False):
    """Compiled by __graft_entry__.build_oracle (content-hashed, one process at a time)."""
    import sys
    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    import __graft_entry__ as graft
    assert graft.ORACLE_LIB == LIB
    return graft.build_oracle(force)


def lib():
    global _lib
    if _lib is None:
        build()
        l = ctypes.CDLL(LIB)
        vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        l.oracle_bbox_from_pixels.argtypes = [vp, vp, f32, i32, i32, vp, vp]
        l.oracle_bbox_from_pixels.restype = None
        l.oracle_grid_params.argtypes = [vp, vp, i32, f32, vp, vp, vp, vp, vp]
        l.oracle_grid_params.restype = None
        l.oracle_tsdf_cal.argtypes = [vp, vp, vp, f32, f32, i32, i32, f32, i32, i32, vp]
        l.oracle_tsdf_cal.restype = None
        l.oracle_dtsdf_batch.argtypes = [i32, vp, vp, vp, i32, i32, i32, vp, f32, f32, i32, i32, vp, vp, vp, i32]
        l.oracle_dtsdf_batch.restype = i32
        l.oracle_dtsdf_aug_batch.argtypes = [i32, vp, vp, vp, vp, i32, i32, f32, f32, i32, i32, vp, vp, vp]
        l.oracle_dtsdf_aug_batch.restype = i32
        _lib = l
    return _lib


def dtsdf_batch(headers, offsets, packed, R=32, layout=0, bbox_mode=0, aux=None, trunc_mul=3.0,
                focal=241.42, cx=160, cy=120, nthreads=0, out=None):
    """Whole path on a packed batch.  Returns (vol, max_l, mid_p, threads_used)."""
    headers = np.ascontiguousarray(headers, dtype=np.int32).reshape(-1, 6)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    packed = np.ascontiguousarray(packed, dtype=np.float32)
    n = headers.shape[0]
    vol = out if out is not None else np.empty((n, 3, R, R, R), dtype=np.float32)
    max_l = np.zeros(n, dtype=np.float32)
    mid_p = np.zeros((n, 3), dtype=np.float32)
    aux_p = None
    if aux is not None:
        aux = np.ascontiguousarray(aux, dtype=np.float32)
        aux_p = aux.ctypes.data
    used = lib().oracle_dtsdf_batch(n, headers.ctypes.data, packed.ctypes.data, offsets.ctypes.data, R, layout,
                                    bbox_mode, aux_p, trunc_mul, focal, cx, cy, vol.ctypes.data,
                                    max_l.ctypes.data, mid_p.ctypes.data, nthreads)
    return vol, max_l, mid_p, used


def dtsdf_aug_batch(headers, offsets, packed, aug, R=32, layout=0, trunc_mul=3.0, focal=241.42, cx=160, cy=120):
    """Augmented D-TSDF (the specification in dtsdf_oracle.c; parity unpinned).
    aug float32[n,21] = A (3x3), Ainv (3x3), c (3).  Returns (vol, max_l, mid_p)."""
    headers = np.ascontiguousarray(headers, dtype=np.int32).reshape(-1, 6)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    packed = np.ascontiguousarray(packed, dtype=np.float32)
    aug = np.ascontiguousarray(aug, dtype=np.float32).reshape(-1, 21)
    n = headers.shape[0]
    vol = np.empty((n, 3, R, R, R), dtype=np.float32)
    max_l = np.zeros(n, dtype=np.float32)
    mid_p = np.zeros((n, 3), dtype=np.float32)
    lib().oracle_dtsdf_aug_batch(n, headers.ctypes.data, packed.ctypes.data, offsets.ctypes.data, aug.ctypes.data,
                                 R, layout, trunc_mul, focal, cx, cy, vol.ctypes.data, max_l.ctypes.data,
                                 mid_p.ctypes.data)
    return vol, max_l, mid_p
