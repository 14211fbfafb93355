"""Host-side API of the B200 D-TSDF voxeliser on top of the C ABI (capi.py).

Two ways in:

* ``Voxelizer`` - NumPy in, NumPy out, through the library's *host* entry points
  (its own streams, chunked copy/compute overlap).  The drop-in modules in ``pre/``
  (process.py, tsdf_for.py, tsdf_numba.py) are thin wrappers over this class.
* ``voxelize_device`` and friends - torch CUDA tensors in and out through the
  *device* entry points on torch's current stream; nothing leaves HBM.  torch is
  only used for device memory and streams.
"""
import ctypes
import functools
import threading

import numpy as np

from . import capi
from .capi import (BBOX_EXTREMA, BBOX_GRID, BBOX_PIXELS, LAYOUT_CXYZ, LAYOUT_CZYX, check,
                   load_library)

FOCAL = 241.42          # pre/process.py:7, pre/tsdf_numba.py:8
CENTER_X = 160          # pre/process.py:133, pre/tsdf_numba.py:9
CENTER_Y = 120          # pre/process.py:134, pre/tsdf_numba.py:10
VOXEL_RES = 32          # pre/process.py:88


def frame_depth(frame):
    """Depth vector of a frame record: key 'depth' (process.py / tsdf_for.py) or 'data'
    (pre/tsdf_numba.py:132-133)."""
    if "depth" in frame:
        return frame["depth"]
    return frame["data"]


def pack_frames(frames, pinned=False, scratch=None):
    """Packs frame records into the batched input format of include/dtsdf_b200.h.

    Returns ``(headers int32[n,6], offsets int64[n], packed float32[packed_len], max_crop_px)``;
    crop starts are 16-byte aligned and the tail padding is zero.  With ``pinned=True`` the
    packed buffer is page-locked and a fifth value, its owning ``PinnedBuffer``, is returned.
    ``scratch``: a callable ``n_floats -> float32 array`` that provides the packed buffer instead
    (``Voxelizer`` hands out its page-locked staging area this way).
    """
    n = len(frames)
    headers = np.empty((n, 6), dtype=np.int32)
    for i, fr in enumerate(frames):
        h = np.asarray(fr["header"]).reshape(-1)
        if h.size < 6:
            raise ValueError("frame %d: header needs 6 int32 values" % i)
        headers[i] = h[:6]
    offsets, packed_len, max_px = capi.pack_plan(headers)
    if scratch is not None:
        buf = None
        packed = scratch(max(packed_len, 4))
        packed[:] = 0
    elif pinned:
        buf = capi.PinnedBuffer(max(packed_len, 4), np.float32)
        packed = buf.array
        packed[:] = 0
    else:
        buf = None
        packed = np.zeros(max(packed_len, 4), dtype=np.float32)
    sizes = (headers[:, 4] - headers[:, 2]).astype(np.int64) * (headers[:, 5] - headers[:, 3])
    depths = [np.asarray(frame_depth(fr)).reshape(-1) for fr in frames]
    for i, d in enumerate(depths):
        if d.size != sizes[i]:
            raise ValueError("frame %d: depth has %d values, bounding box has %d pixels" % (i, d.size, sizes[i]))
    if n and not (sizes & 3).any() and all(d.dtype == np.float32 for d in depths):
        # no padding between the crops (every crop a multiple of four pixels): one concatenation
        np.concatenate(depths, out=packed[:int(sizes.sum())])
    else:
        for i, d in enumerate(depths):
            packed[offsets[i]:offsets[i] + d.size] = d
    if pinned:
        return headers, offsets, packed, max_px, buf
    return headers, offsets, packed, max_px


def _serialised(method):
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with self._lock:
            return method(self, *args, **kwargs)
    return wrapper


class Voxelizer:
    """NumPy-facing voxeliser bound to one CUDA device (host entry points of the C ABI).  Calls on one
    instance are serialised by a lock; the process-wide ``default_voxelizer()`` is therefore safe to share
    between threads, and no call depends on state left by an earlier one unless its docstring says so."""

    def __init__(self, device=0, voxel_res=VOXEL_RES, focal=FOCAL, center_x=CENTER_X, center_y=CENTER_Y,
                 trunc_mul=3.0):
        self._lib = load_library()
        self.voxel_res = int(voxel_res)
        self.focal = float(focal)
        self.center_x = int(center_x)
        self.center_y = int(center_y)
        self.trunc_mul = float(trunc_mul)
        self._ctx = ctypes.c_void_p()
        # one call at a time per context (ctypes drops the GIL during a call; the context's streams,
        # staging buffers and resident clouds are not re-entrant).  Threads that want to overlap use
        # a Voxelizer each.
        self._lock = threading.RLock()
        check(self._lib.dtsdf_ctx_create(int(device), ctypes.byref(self._ctx)))
        # one process per GPU (torchrun): the ranks of a node share its host cores, so each takes its
        # share of them for the expansion of packed volumes
        import os
        try:
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        except ValueError:
            local_world = 1
        if local_world > 1:
            self.set_option(capi.OPT_HOST_THREADS, max(1, min(8, (os.cpu_count() or 1) // local_world)))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.dtsdf_ctx_destroy(self._ctx)
            self._ctx = None
        if getattr(self, "_stage", None) is not None:
            self._stage.free()
            self._stage = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _staging(self, n_floats):
        """Page-locked packing area of the frame-level calls, grown when needed and kept: a host -> device copy
        out of pageable memory is staged by the driver on the calling thread (the process() host call over
        4,000 frames: 37 ms from a NumPy array, 14 ms from page-locked memory), and page-locking memory per call
        costs more than it saves.  Valid until the next frame-level call on this instance."""
        if getattr(self, "_stage", None) is None or self._stage.array.size < n_floats:
            if getattr(self, "_stage", None) is not None:
                self._stage.free()
            self._stage = capi.PinnedBuffer((max(int(n_floats * 1.25), 1 << 20),), np.float32)
        return self._stage.array[:n_floats]

    @_serialised
    def set_option(self, option, value):
        """``dtsdf_ctx_set_option``: capi.OPT_PACKED_D2H (packed device->host transfer of the volumes,
        default on) and capi.OPT_HOST_THREADS (host threads that expand it)."""
        check(self._lib.dtsdf_ctx_set_option(self._ctx, int(option), int(value)))

    @_serialised
    def stat(self, which=capi.STAT_D2H_BYTES):
        """``dtsdf_ctx_get_stat`` of the last ``voxelize_packed`` call."""
        v = ctypes.c_int64(0)
        check(self._lib.dtsdf_ctx_get_stat(self._ctx, int(which), ctypes.byref(v)))
        return int(v.value)

    def config(self, layout=LAYOUT_CXYZ, bbox_mode=BBOX_PIXELS, voxel_res=None):
        cfg = capi.default_config(voxel_res or self.voxel_res, layout, bbox_mode)
        cfg.focal = self.focal
        cfg.center_x = self.center_x
        cfg.center_y = self.center_y
        cfg.trunc_mul = self.trunc_mul
        return cfg

    # ------------------------------------------------------------------ voxelisation
    @_serialised
    def voxelize_packed(self, headers, offsets, packed, layout=LAYOUT_CXYZ, bbox_mode=BBOX_PIXELS,
                        aux=None, voxel_res=None, out_vol=None, out_max_l=None, out_mid_p=None):
        """``dtsdf_voxelize_host`` on an already packed batch (see pack_frames).

        ``out_*`` may be preallocated (e.g. pinned) arrays.  Returns (vol, max_l, mid_p).
        """
        R = int(voxel_res or self.voxel_res)
        headers = np.ascontiguousarray(headers, dtype=np.int32).reshape(-1, 6)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = headers.shape[0]
        assert packed.dtype == np.float32 and packed.flags["C_CONTIGUOUS"]
        if out_vol is None:
            out_vol = np.empty((n, 3, R, R, R), dtype=np.float32)
        if out_max_l is None:
            out_max_l = np.empty(n, dtype=np.float32)
        if out_mid_p is None:
            out_mid_p = np.empty((n, 3), dtype=np.float32)
        if aux is not None:
            aux = np.ascontiguousarray(aux, dtype=np.float32)
            width = 6 if bbox_mode == BBOX_EXTREMA else 5
            if aux.shape != (n, width):
                raise ValueError("aux must be float32[%d,%d]" % (n, width))
        cfg = self.config(layout, bbox_mode, R)
        check(self._lib.dtsdf_voxelize_host(
            self._ctx, ctypes.byref(cfg), n, capi._ptr(packed), int(packed.size), capi._ptr(offsets),
            capi._ptr(headers), capi._ptr(aux), capi._ptr(out_vol), capi._ptr(out_max_l),
            capi._ptr(out_mid_p)))
        return out_vol, out_max_l, out_mid_p

    @_serialised
    def voxelize_packed_lazy(self, headers, offsets, packed, layout=LAYOUT_CXYZ, bbox_mode=BBOX_PIXELS,
                             aux=None, voxel_res=None):
        """``dtsdf_voxelize_host_packed``: like ``voxelize_packed`` but the volumes stay packed in host memory
        (``PackedHostVolumes``, expanded on demand).  Returns (packed_volumes, max_l, mid_p)."""
        R = int(voxel_res or self.voxel_res)
        headers = np.ascontiguousarray(headers, dtype=np.int32).reshape(-1, 6)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = headers.shape[0]
        assert packed.dtype == np.float32 and packed.flags["C_CONTIGUOUS"]
        out_max_l = np.empty(n, dtype=np.float32)
        out_mid_p = np.empty((n, 3), dtype=np.float32)
        if aux is not None:
            aux = np.ascontiguousarray(aux, dtype=np.float32)
        cfg = self.config(layout, bbox_mode, R)
        handle = ctypes.c_void_p()
        check(self._lib.dtsdf_voxelize_host_packed(
            self._ctx, ctypes.byref(cfg), n, capi._ptr(packed), int(packed.size), capi._ptr(offsets),
            capi._ptr(headers), capi._ptr(aux), ctypes.byref(handle), capi._ptr(out_max_l), capi._ptr(out_mid_p)))
        if not handle.value:                       # n == 0
            return None, out_max_l, out_mid_p
        return PackedHostVolumes(handle, self._lib), out_max_l, out_mid_p

    def voxelize_frames_lazy(self, frames, layout=LAYOUT_CXYZ, bbox_mode=BBOX_PIXELS, aux=None, voxel_res=None):
        """Frame records -> (PackedHostVolumes, max_l, mid_p): ``voxelize_packed_lazy`` through the staging area."""
        with self._lock:
            headers, offsets, packed, _ = pack_frames(frames, scratch=self._staging)
            return self.voxelize_packed_lazy(headers, offsets, packed, layout, bbox_mode, aux, voxel_res)

    @_serialised
    def voxelize_frames(self, frames, layout=LAYOUT_CXYZ, bbox_mode=BBOX_PIXELS, aux=None, voxel_res=None):
        """Frame records -> (vol float32[n,3,R,R,R], max_l float32[n], mid_p float32[n,3])."""
        headers, offsets, packed, _ = pack_frames(frames, scratch=self._staging)
        return self.voxelize_packed(headers, offsets, packed, layout, bbox_mode, aux, voxel_res)

    # ------------------------------------------------------------------ process.py pieces
    @_serialised
    def point_cloud(self, frame):
        """DataProcess.point_cloud (pre/process.py:30-68): float64[n_valid,3]; the cloud also
        stays resident on the device for ``resample_extrema``."""
        header = np.ascontiguousarray(np.asarray(frame["header"]).reshape(-1)[:6], dtype=np.int32)
        depth = np.ascontiguousarray(frame_depth(frame), dtype=np.float32).reshape(-1)
        n_px = int(header[4] - header[2]) * int(header[5] - header[3])
        if depth.size != n_px:
            raise ValueError("depth has %d values, bounding box has %d pixels" % (depth.size, n_px))
        pts = np.empty((max(n_px, 1), 3), dtype=np.float64)
        n_valid = ctypes.c_int32(0)
        check(self._lib.dtsdf_point_cloud_host(self._ctx, ctypes.c_double(self.focal), capi._ptr(header),
                                               capi._ptr(depth), capi._ptr(pts), ctypes.byref(n_valid)))
        return pts[:n_valid.value].copy()

    @_serialised
    def resample_extrema(self, point_num, indices=None, points=None, want_points=True):
        """Gather ``point_num`` rows by ``indices`` (pre/process.py:82) from ``points`` (or the cloud
        left on the device by ``point_cloud``) and reduce (max, min) (pre/process.py:102-120).

        Returns ``(pc float64[point_num,3] or None, point_max float32[3], point_min float32[3])``.
        """
        point_num = int(point_num)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
        if idx is not None and idx.size != point_num:
            raise ValueError("indices must have point_num entries")
        pts = None
        n_points = 0
        if points is not None:
            pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
            n_points = pts.shape[0]
        out_pc = np.empty((point_num, 3), dtype=np.float64) if want_points else None
        ext = np.empty(6, dtype=np.float32)
        check(self._lib.dtsdf_resample_extrema_host(self._ctx, point_num, capi._ptr(pts), n_points,
                                                    capi._ptr(idx), capi._ptr(out_pc), capi._ptr(ext)))
        return out_pc, ext[:3].copy(), ext[3:].copy()

    def process_frames(self, frames, point_num=6000, want_points=True, layout=LAYOUT_CXYZ, voxel_res=None):
        """``DataProcess(frame, gt, point_num).process()`` (pre/process.py:13-28) for a list of frame
        records: ``pack_frames`` + ``process_packed``."""
        with self._lock:
            headers, offsets, packed, _ = pack_frames(frames, scratch=self._staging)
            return self.process_packed(headers, offsets, packed, point_num, want_points, layout, voxel_res)

    @_serialised
    def process_packed(self, headers, offsets, packed, point_num=6000, want_points=True, layout=LAYOUT_CXYZ,
                       voxel_res=None, out_vol=None):
        """``DataProcess.process()`` for a packed batch in ONE library call (``dtsdf_process_host``): point
        clouds on the device, the resampling index vectors drawn frame by frame, in order, from the *global*
        ``np.random`` stream - by the library, in C, with NumPy's own algorithm (MT19937 + masked
        rejection; the stream is handed over with ``np.random.get_state()`` and given back advanced) - gather,
        extrema, voxel loop on the GPU.  A seeded run therefore reproduces the reference's frame loop
        (pre/read_MSRA.py:98-111) bit for bit, and whatever the caller draws afterwards too.

        Returns ``(point_clouds float64[n,point_num,3] or None, tsdf float32[n,3,R,R,R],
        max_l float32[n], mid_p float32[n,3])``.  A frame without a single point raises ValueError like the
        reference's ``np.random.randint(0, 0)``.
        """
        R = int(voxel_res or self.voxel_res)
        point_num = int(point_num)
        headers = np.ascontiguousarray(headers, dtype=np.int32).reshape(-1, 6)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        assert packed.dtype == np.float32 and packed.flags["C_CONTIGUOUS"]
        n = headers.shape[0]
        pcs = np.empty((n, point_num, 3), dtype=np.float64) if want_points else None
        vol = np.empty((n, 3, R, R, R), dtype=np.float32) if out_vol is None else out_vol
        max_l = np.empty(n, dtype=np.float32)
        mid_p = np.empty((n, 3), dtype=np.float32)
        cfg = self.config(layout, BBOX_EXTREMA, R)
        with capi.NumpyGlobalStream() as st:
            check(self._lib.dtsdf_process_host(
                self._ctx, ctypes.byref(cfg), ctypes.c_double(self.focal), n, capi._ptr(packed), int(packed.size),
                capi._ptr(offsets), capi._ptr(headers), point_num, st.key_ptr, st.pos_ref, capi._ptr(pcs),
                capi._ptr(vol), capi._ptr(max_l), capi._ptr(mid_p)))
        return pcs, vol, max_l, mid_p

    def process_frames_two_calls(self, frames, point_num=6000, want_points=True, layout=LAYOUT_CXYZ, voxel_res=None):
        """The same result through ``dtsdf_process_begin_host`` / ``dtsdf_process_finish_host`` with the index
        vectors drawn HERE by ``np.random.randint``, one call per frame as the reference does
        (pre/process.py:75-80): the form that depends on nothing but NumPy's public API (``process_packed``
        is checked against it)."""
        point_num = int(point_num)
        with self._lock:
            n_valid = self.process_begin(frames)
            n = n_valid.size
            idx = np.empty((n, point_num), dtype=np.int32)
            for i in range(n):
                point_shape = int(n_valid[i])
                if point_shape < point_num:
                    idx[i] = np.arange(0, point_num, 1, dtype=np.int32)
                    idx[i, point_shape:] = np.random.randint(0, point_shape, size=point_num - point_shape)
                else:
                    idx[i] = np.random.randint(0, point_shape, size=point_num)
            return self.process_finish(idx, point_num, want_points, layout, voxel_res)

    @_serialised
    def process_begin(self, frames):
        """``dtsdf_process_begin_host``: uploads the frames, back-projects and compacts every point cloud
        (pre/process.py:30-68; crops and clouds stay resident in the context until ``process_finish``) and
        returns the number of points per frame, int32[n] - what the caller needs to draw the resampling
        indices from its own random stream."""
        headers, offsets, packed, _ = pack_frames(frames)
        n = headers.shape[0]
        n_valid = np.zeros(n, dtype=np.int32)
        check(self._lib.dtsdf_process_begin_host(self._ctx, ctypes.c_double(self.focal), n, capi._ptr(packed),
                                                 int(packed.size), capi._ptr(offsets), capi._ptr(headers),
                                                 capi._ptr(n_valid)))
        return n_valid

    @_serialised
    def process_finish(self, indices, point_num=6000, want_points=True, layout=LAYOUT_CXYZ, voxel_res=None):
        """``dtsdf_process_finish_host`` on the frames of the last ``process_begin``: gather by ``indices``
        int32[n, point_num] (pre/process.py:82), extrema (:102-120), voxel loop.  No other call on this
        Voxelizer may come in between.  Returns ``(point_clouds or None, tsdf, max_l, mid_p)``."""
        R = int(voxel_res or self.voxel_res)
        point_num = int(point_num)
        idx = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, point_num)
        n = idx.shape[0]
        pcs = np.empty((n, point_num, 3), dtype=np.float64) if want_points else None
        vol = np.empty((n, 3, R, R, R), dtype=np.float32)
        max_l = np.empty(n, dtype=np.float32)
        mid_p = np.empty((n, 3), dtype=np.float32)
        cfg = self.config(layout, BBOX_EXTREMA, R)
        check(self._lib.dtsdf_process_finish_host(self._ctx, ctypes.byref(cfg), point_num, capi._ptr(idx),
                                                  capi._ptr(pcs), capi._ptr(vol), capi._ptr(max_l),
                                                  capi._ptr(mid_p)))
        return pcs, vol, max_l, mid_p

    @_serialised
    def normalize_joints(self, joints, max_l, mid_p, clamp01=False):
        """(gt - mid_p)/max_l + 0.5 (pre/joint_nor.py:8-18), float32[n,63]."""
        j = np.ascontiguousarray(joints, dtype=np.float32).reshape(-1, 63)
        n = j.shape[0]
        ml = np.ascontiguousarray(max_l, dtype=np.float32).reshape(n)
        mp = np.ascontiguousarray(mid_p, dtype=np.float32).reshape(n, 3)
        out = np.empty_like(j)
        check(self._lib.dtsdf_normalize_joints_host(self._ctx, n, capi._ptr(j), capi._ptr(ml), capi._ptr(mp),
                                                    1 if clamp01 else 0, capi._ptr(out)))
        return out


class PackedHostVolumes:
    """Volumes left packed in host memory by ``Voxelizer.voxelize_packed_lazy`` (``dtsdf_voxelize_host_packed``):
    ~47 KB per volume instead of 393 KB.  Sequence-like: ``len(p)``, ``p[i]`` -> float32[3,R,R,R], ``p[a:b]`` and
    ``p[index array]`` -> float32[k,3,R,R,R], expanded on demand by ``dtsdf_expand_frames``; ``shape`` and
    ``dtype`` as the dense array would have them, so a ``torch.utils.data.Dataset`` can hold it where it held
    the dense array (3D_CNN/dataset.py:35,73-79)."""

    dtype = np.dtype(np.float32)

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib
        n, r, b = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int64(0)
        check(lib.dtsdf_packed_host_info(handle, ctypes.byref(n), ctypes.byref(r), ctypes.byref(b)))
        self.n_frames, self.voxel_res, self.nbytes = int(n.value), int(r.value), int(b.value)

    @property
    def shape(self):
        R = self.voxel_res
        return (self.n_frames, 3, R, R, R)

    def __len__(self):
        return self.n_frames

    def expand(self, first=0, count=None, out=None, threads=0):
        """Frames [first, first+count) as float32[count,3,R,R,R] (into ``out`` if given)."""
        count = self.n_frames - first if count is None else int(count)
        R = self.voxel_res
        if out is None:
            out = np.empty((count, 3, R, R, R), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.size == count * 3 * R ** 3
        check(self._lib.dtsdf_expand_frames(self._h, int(first), count, capi._ptr(out), int(threads)))
        return out

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            i = int(index)
            if i < 0:
                i += self.n_frames
            if not 0 <= i < self.n_frames:
                raise IndexError(index)
            return self.expand(i, 1, threads=1)[0]
        if isinstance(index, slice):
            a, b, st = index.indices(self.n_frames)
            if st == 1:
                return self.expand(a, max(b - a, 0))
            index = np.arange(a, b, st)
        idx = np.asarray(index).reshape(-1)
        R = self.voxel_res
        out = np.empty((idx.size, 3, R, R, R), dtype=np.float32)
        for k, i in enumerate(idx):
            self.expand(int(i) % self.n_frames if i < 0 else int(i), 1, out=out[k:k + 1], threads=1)
        return out

    def close(self):
        if self._h is not None:
            self._lib.dtsdf_packed_host_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PackedHostVolumeList:
    """Several ``PackedHostVolumes`` (e.g. one per gesture) behind one index space."""

    dtype = np.dtype(np.float32)

    def __init__(self, parts):
        self.parts = list(parts)
        self.starts = np.concatenate([[0], np.cumsum([len(p) for p in self.parts])]).astype(np.int64)
        self.voxel_res = self.parts[0].voxel_res if self.parts else 32
        self.nbytes = sum(p.nbytes for p in self.parts)

    @property
    def shape(self):
        R = self.voxel_res
        return (int(self.starts[-1]), 3, R, R, R)

    def __len__(self):
        return int(self.starts[-1])

    def _one(self, i):
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        k = int(np.searchsorted(self.starts, i, side="right")) - 1
        return self.parts[k][i - int(self.starts[k])]

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            return self._one(index)
        if isinstance(index, slice):
            index = np.arange(*index.indices(len(self)))
        idx = np.asarray(index).reshape(-1)
        R = self.voxel_res
        out = np.empty((idx.size, 3, R, R, R), dtype=np.float32)
        for k, i in enumerate(idx):
            out[k] = self._one(i)
        return out


_default_voxelizers = {}


def default_voxelizer(device=0):
    """Process-wide Voxelizer used by the drop-in modules in ``pre/``."""
    v = _default_voxelizers.get(device)
    if v is None:
        v = Voxelizer(device=device)
        _default_voxelizers[device] = v
    return v


# ---------------------------------------------------------------------- device-resident path
def _torch_stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def voxelize_device(depth_packed, offsets, headers, max_crop_px=0, cfg=None, aux=None,
                    out_vol=None, out_max_l=None, out_mid_p=None, stream=None, joints=None, joint_flags=0,
                    out_joints=None):
    """``dtsdf_voxelize_batch`` on torch CUDA tensors, asynchronous on torch's current stream.

    depth_packed float32[packed_len], offsets int64[n], headers int32[n,6] - all on the same
    CUDA device.  Returns (vol float32[n,3,R,R,R], max_l float32[n], mid_p float32[n,3]).
    With ``joints`` (float32[n,63]) the joint normalisation runs in the same launch
    (``dtsdf_voxelize_joints_batch``) and a fourth value, float32[n,63], is returned.
    """
    import torch
    if not (depth_packed.is_cuda and offsets.is_cuda and headers.is_cuda):
        raise RuntimeError("voxelize_device needs CUDA tensors; there is no CPU fallback")
    assert depth_packed.dtype == torch.float32 and offsets.dtype == torch.int64 and headers.dtype == torch.int32
    assert depth_packed.is_contiguous() and offsets.is_contiguous() and headers.is_contiguous()
    cfg = cfg or capi.default_config()
    n = int(headers.shape[0])
    R = int(cfg.voxel_res)
    dev = depth_packed.device
    if out_vol is None:
        out_vol = torch.empty((n, 3, R, R, R), dtype=torch.float32, device=dev)
    if out_max_l is None:
        out_max_l = torch.empty((n,), dtype=torch.float32, device=dev)
    if out_mid_p is None:
        out_mid_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    aux_ptr = None
    if aux is not None:
        assert aux.is_cuda and aux.dtype == torch.float32 and aux.is_contiguous()
        aux_ptr = aux.data_ptr()
    if joints is not None:
        assert joints.is_cuda and joints.dtype == torch.float32 and joints.is_contiguous() and joints.numel() == n * 63
        if out_joints is None:
            out_joints = torch.empty((n, 63), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(load_library().dtsdf_voxelize_joints_batch(
                ctypes.byref(cfg), n, depth_packed.data_ptr(), offsets.data_ptr(), headers.data_ptr(), aux_ptr,
                int(max_crop_px), joints.data_ptr(), int(joint_flags), out_vol.data_ptr(), out_max_l.data_ptr(),
                out_mid_p.data_ptr(), out_joints.data_ptr(), _torch_stream_ptr(stream)))
        return out_vol, out_max_l, out_mid_p, out_joints
    with torch.cuda.device(dev):
        check(load_library().dtsdf_voxelize_batch(
            ctypes.byref(cfg), n, depth_packed.data_ptr(), offsets.data_ptr(), headers.data_ptr(), aux_ptr,
            int(max_crop_px), out_vol.data_ptr(), out_max_l.data_ptr(), out_mid_p.data_ptr(),
            _torch_stream_ptr(stream)))
    return out_vol, out_max_l, out_mid_p


def count_valid_device(depth_packed, offsets, headers, out=None, stream=None):
    """``dtsdf_count_valid_batch``: int32[n] pixels with d != 0 per frame (``hand_points.shape[0]`` of
    pre/process.py:70-73), asynchronous on torch's current stream."""
    import torch
    assert depth_packed.is_cuda and offsets.dtype == torch.int64 and headers.dtype == torch.int32
    n = int(headers.shape[0])
    if out is None:
        out = torch.empty((n,), dtype=torch.int32, device=depth_packed.device)
    with torch.cuda.device(depth_packed.device):
        check(load_library().dtsdf_count_valid_batch(n, depth_packed.data_ptr(), offsets.data_ptr(), headers.data_ptr(),
                                                     out.data_ptr(), _torch_stream_ptr(stream)))
    return out


def voxelize_sampled_device(depth_packed, offsets, headers, sample_idx, max_crop_px=0, cfg=None, focal=FOCAL,
                            out_vol=None, out_max_l=None, out_mid_p=None, stream=None):
    """``dtsdf_voxelize_sampled_batch``: ``DataProcess.process()`` (pre/process.py:13-28) for a batch in one launch,
    without the clouds.  ``sample_idx`` int32[n, point_num]: every frame's ``rand_idx`` (pre/process.py:75-80).
    Returns (vol, max_l, mid_p) like ``voxelize_device``."""
    import torch
    if not (depth_packed.is_cuda and offsets.is_cuda and headers.is_cuda and sample_idx.is_cuda):
        raise RuntimeError("voxelize_sampled_device needs CUDA tensors; there is no CPU fallback")
    assert depth_packed.dtype == torch.float32 and offsets.dtype == torch.int64 and headers.dtype == torch.int32
    assert sample_idx.dtype == torch.int32 and sample_idx.is_contiguous() and sample_idx.dim() == 2
    cfg = cfg or capi.default_config()
    n = int(headers.shape[0])
    assert int(sample_idx.shape[0]) == n
    R = int(cfg.voxel_res)
    dev = depth_packed.device
    if out_vol is None:
        out_vol = torch.empty((n, 3, R, R, R), dtype=torch.float32, device=dev)
    if out_max_l is None:
        out_max_l = torch.empty((n,), dtype=torch.float32, device=dev)
    if out_mid_p is None:
        out_mid_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(load_library().dtsdf_voxelize_sampled_batch(
            ctypes.byref(cfg), float(focal), n, depth_packed.data_ptr(), offsets.data_ptr(), headers.data_ptr(),
            sample_idx.data_ptr(), int(sample_idx.shape[1]), int(max_crop_px), out_vol.data_ptr(), out_max_l.data_ptr(),
            out_mid_p.data_ptr(), _torch_stream_ptr(stream)))
    return out_vol, out_max_l, out_mid_p


def normalize_joints_device(joints, max_l, mid_p, clamp01=True, out=None, stream=None):
    """``dtsdf_normalize_joints_batch`` on torch CUDA tensors (3D_CNN/train.py:236-244)."""
    import torch
    assert joints.is_cuda and joints.dtype == torch.float32 and joints.is_contiguous()
    n = int(joints.shape[0])
    if out is None:
        out = torch.empty((n, 63), dtype=torch.float32, device=joints.device)
    with torch.cuda.device(joints.device):
        check(load_library().dtsdf_normalize_joints_batch(
            n, joints.data_ptr(), max_l.contiguous().data_ptr(), mid_p.contiguous().data_ptr(),
            1 if clamp01 else 0, out.data_ptr(), _torch_stream_ptr(stream)))
    return out


def shard_range(n_frames, rank, world_size):
    """Contiguous frame range [lo, hi) of ``rank``: frames are independent, so
    the batch is cut into ``world_size`` near-equal contiguous shards and nothing is exchanged
    while voxelising."""
    base, rem = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_packed(headers, offsets, packed, rank, world_size):
    """Rank-local view of a packed batch: (headers, rebased offsets, packed slice)."""
    n = headers.shape[0]
    lo, hi = shard_range(n, rank, world_size)
    if hi == lo:
        return headers[lo:hi], offsets[lo:hi], packed[:0]
    e0 = int(offsets[lo])
    h = headers[hi - 1]
    n_last = int(h[4] - h[2]) * int(h[5] - h[3])
    e1 = int(offsets[hi - 1]) + ((n_last + 3) & ~3)
    return headers[lo:hi], offsets[lo:hi] - e0, packed[e0:e1]
