"""ctypes binding of the C ABI in include/dtsdf_b200.h.

The shared library ``lib/libdtsdf_b200.so`` is built in-tree by
``__graft_entry__.build()``.  There is no fallback: if the library is missing, or no
CUDA device is usable, the calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdtsdf_b200.so")

OK = 0
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED = -2
ERR_CUDA = -3
ERR_NO_DEVICE = -4
ERR_ALLOC = -5
ERR_BAD_HEADER = -6
ERR_INTERNAL = -7
ERR_EMPTY_CLOUD = -8

LAYOUT_CXYZ = 0
LAYOUT_CZYX = 1
BBOX_PIXELS = 0
BBOX_EXTREMA = 1
BBOX_GRID = 2
OPT_PACKED_D2H = 1
OPT_HOST_THREADS = 2
OPT_CHUNK_MB = 3
OPT_LITERAL_PUSH = 4
OPT_TEST_FAIL_CHUNK = 100
STAT_D2H_BYTES = 1
JOINT_CLAMP01 = 1
JOINT_FLIP_Z = 2


class DtsdfError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = "dtsdf error %d" % code
        if detail:
            msg += ": " + detail
        super().__init__(msg)


class Config(ctypes.Structure):
    """``dtsdf_config`` (include/dtsdf_b200.h)."""

    _fields_ = [
        ("voxel_res", ctypes.c_int32),
        ("layout", ctypes.c_int32),
        ("bbox_mode", ctypes.c_int32),
        ("center_x", ctypes.c_int32),
        ("center_y", ctypes.c_int32),
        ("focal", ctypes.c_float),
        ("trunc_mul", ctypes.c_float),
    ]


_lib = None

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_cfgp = ctypes.POINTER(Config)

_SIGNATURES = {
    "dtsdf_abi_version": (ctypes.c_int, []),
    "dtsdf_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "dtsdf_last_cuda_error": (ctypes.c_char_p, []),
    "dtsdf_default_config": (None, [_cfgp]),
    "dtsdf_device_count": (ctypes.c_int, []),
    "dtsdf_launch_count": (_i64, []),
    "dtsdf_enable_peer_access": (ctypes.c_int, [_i32, _i32]),
    "dtsdf_shared_alloc": (ctypes.c_int, [_i32, _i64, ctypes.POINTER(_vp), _vp]),
    "dtsdf_shared_free": (ctypes.c_int, [_i32, _vp]),
    "dtsdf_shared_open": (ctypes.c_int, [_i32, _vp, ctypes.POINTER(_vp)]),
    "dtsdf_shared_close": (ctypes.c_int, [_i32, _vp]),
    "dtsdf_pack_plan": (ctypes.c_int, [_i32, _vp, _vp, ctypes.POINTER(_i64), ctypes.POINTER(_i32)]),
    "dtsdf_voxelize_batch": (ctypes.c_int, [_cfgp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dtsdf_voxelize_joints_batch": (ctypes.c_int, [_cfgp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp,
                                                   _vp]),
    "dtsdf_voxelize_aug_batch": (ctypes.c_int, [_cfgp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp,
                                                _vp]),
    "dtsdf_count_valid_batch": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_voxelize_sampled_batch": (ctypes.c_int, [_cfgp, ctypes.c_double, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp,
                                                    _vp]),
    "dtsdf_point_cloud_batch": (ctypes.c_int, [_i32, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_resample_extrema_batch": (ctypes.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_sample_cloud_batch": (ctypes.c_int, [_i32, ctypes.c_double, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "dtsdf_normalize_joints_batch": (ctypes.c_int, [_i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "dtsdf_pack_volumes_batch": (ctypes.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_unpack_volumes_batch": (ctypes.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_ctx_create": (ctypes.c_int, [_i32, ctypes.POINTER(_vp)]),
    "dtsdf_ctx_destroy": (None, [_vp]),
    "dtsdf_ctx_set_option": (ctypes.c_int, [_vp, _i32, _i64]),
    "dtsdf_ctx_get_stat": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_i64)]),
    "dtsdf_host_alloc": (ctypes.c_int, [ctypes.POINTER(_vp), _i64]),
    "dtsdf_host_free": (None, [_vp]),
    "dtsdf_voxelize_host": (ctypes.c_int, [_vp, _cfgp, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_voxelize_host_packed": (ctypes.c_int, [_vp, _cfgp, _i32, _vp, _i64, _vp, _vp, _vp, ctypes.POINTER(_vp), _vp, _vp]),
    "dtsdf_expand_frames": (ctypes.c_int, [_vp, _i32, _i32, _vp, _i32]),
    "dtsdf_packed_host_info": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), ctypes.POINTER(_i32), ctypes.POINTER(_i64)]),
    "dtsdf_packed_host_free": (None, [_vp]),
    "dtsdf_point_cloud_host": (ctypes.c_int, [_vp, ctypes.c_double, _vp, _vp, _vp, ctypes.POINTER(_i32)]),
    "dtsdf_resample_extrema_host": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dtsdf_process_begin_host": (ctypes.c_int, [_vp, ctypes.c_double, _i32, _vp, _i64, _vp, _vp, _vp]),
    "dtsdf_process_finish_host": (ctypes.c_int, [_vp, _cfgp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "dtsdf_normalize_joints_host": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "dtsdf_np_randint": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), _i32, _i32, _vp]),
    "dtsdf_resample_indices": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), _i32, _vp, _i32, _vp]),
    "dtsdf_process_host": (ctypes.c_int, [_vp, _cfgp, ctypes.c_double, _i32, _vp, _i64, _vp, _vp, _i32, _vp,
                                          ctypes.POINTER(_i32), _vp, _vp, _vp, _vp]),
}


def load_library():
    """Loads (once) and returns the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dtsdf_abi_version() != 1:
        raise RuntimeError("libdtsdf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc == OK:
        return
    if rc == ERR_EMPTY_CLOUD:
        raise ValueError("low >= high")          # what the reference's np.random.randint(0, 0) raises
    lib = load_library()
    detail = lib.dtsdf_error_string(rc).decode()
    if rc in (ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC, ERR_INTERNAL):
        extra = lib.dtsdf_last_cuda_error().decode()
        if extra:
            detail += " [" + extra + "]"
    raise DtsdfError(rc, detail)


def default_config(voxel_res=32, layout=LAYOUT_CXYZ, bbox_mode=BBOX_PIXELS):
    cfg = Config()
    load_library().dtsdf_default_config(ctypes.byref(cfg))
    cfg.voxel_res = voxel_res
    cfg.layout = layout
    cfg.bbox_mode = bbox_mode
    return cfg


def device_count():
    return int(load_library().dtsdf_device_count())


def launch_count():
    return int(load_library().dtsdf_launch_count())


def _ptr(a):
    """Host pointer of a C-contiguous NumPy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def pack_plan(headers):
    """(crop_offsets int64[n], packed_len, max_crop_px) for int32 headers [n,6]."""
    headers = np.ascontiguousarray(headers, dtype=np.int32).reshape(-1, 6)
    n = headers.shape[0]
    offsets = np.empty(n, dtype=np.int64)
    packed_len = _i64(0)
    max_px = _i32(0)
    check(load_library().dtsdf_pack_plan(n, _ptr(headers), _ptr(offsets), ctypes.byref(packed_len),
                                         ctypes.byref(max_px)))
    return offsets, int(packed_len.value), int(max_px.value)


class PinnedBuffer:
    """Page-locked host memory from dtsdf_host_alloc, viewed as a NumPy array."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = _vp()
        check(load_library().dtsdf_host_alloc(ctypes.byref(self._p), max(nbytes, 1)))
        buf = (ctypes.c_byte * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None and self._p.value:
            self.array = None
            load_library().dtsdf_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class NumpyGlobalStream:
    """The caller's global ``np.random`` stream (legacy MT19937) handed to the library and back:

        with NumpyGlobalStream() as st:
            lib.dtsdf_...(st.key_ptr, st.pos_ref, ...)

    On exit - also when the call raised - ``np.random.set_state`` receives the advanced state, so what
    follows in the caller's program draws what it would have drawn after the reference's own
    ``np.random.randint`` calls (pre/process.py:75-80)."""

    def __enter__(self):
        import numpy as _np
        self._st = _np.random.get_state()
        if self._st[0] != "MT19937":
            raise RuntimeError("np.random is not the legacy MT19937 stream")
        self.key = _np.ascontiguousarray(self._st[1], dtype=_np.uint32).copy()
        self.pos = _i32(int(self._st[2]))
        self.key_ptr = self.key.ctypes.data
        self.pos_ref = ctypes.byref(self.pos)
        return self

    def __exit__(self, *exc):
        import numpy as _np
        _np.random.set_state((self._st[0], self.key, int(self.pos.value), self._st[3], self._st[4]))
        return False


def np_randint(high, count):
    """``np.random.randint(0, high, size=count)`` drawn by the library from the global stream (int32)."""
    out = np.empty(int(count), dtype=np.int32)
    with NumpyGlobalStream() as st:
        check(load_library().dtsdf_np_randint(st.key_ptr, st.pos_ref, int(high), int(count), _ptr(out)))
    return out


def resample_indices(n_valid, point_num):
    """The index vectors of ``DataProcess.set_length`` (pre/process.py:70-84) for frames with ``n_valid``
    points each, in order, from the global ``np.random`` stream: int32[n, point_num]."""
    n_valid = np.ascontiguousarray(n_valid, dtype=np.int32).reshape(-1)
    out = np.empty((n_valid.size, int(point_num)), dtype=np.int32)
    with NumpyGlobalStream() as st:
        check(load_library().dtsdf_resample_indices(st.key_ptr, st.pos_ref, n_valid.size, _ptr(n_valid),
                                                    int(point_num), _ptr(out)))
    return out
