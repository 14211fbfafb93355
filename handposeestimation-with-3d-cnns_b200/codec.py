"""Lossless packed form of D-TSDF volumes on the device (``dtsdf_pack_volumes_batch`` /
``dtsdf_unpack_volumes_batch`` of the C ABI): 2-bit codes for voxels that are 0 / +1 / -1 in all three
channels, literal (x, y, z) triples for the rest - about 47 KB in use per 393 KB volume at R = 32.

The reference has nothing like it (single process, volumes go to disk as float64,
pre/read_MSRA.py:119-126); it exists for the one collective of this path, the optional gather of the
volumes of all ranks into a training rank (``gather.PackedVolumeGather``).
torch is used for device memory and streams only.
"""
import ctypes

import torch

from .capi import check, load_library


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


class PackedVolumes:
    """The three arrays of the packed form for up to ``n_frames`` volumes of ``voxel_res``^3 voxels:
    ``codes`` uint8[n, R^3/4], ``lit_count`` int32[n] (uint32 on the device), ``lits`` float32[n, 3*R^3]
    (only the first ``3*lit_count[f]`` floats of a row are ever written or read).  ``storage``: an optional
    uint8 tensor to carve the arrays from (e.g. memory shared with other processes)."""

    def __init__(self, n_frames, voxel_res=32, device="cuda", storage=None):
        n, R = int(n_frames), int(voxel_res)
        r3 = R * R * R
        self.n_frames, self.voxel_res = n, R
        sizes = [n * (r3 // 4), n * 4, n * 3 * r3 * 4]
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 255) & ~255
        self.nbytes = max(total, 256)
        if storage is None:
            storage = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        assert storage.dtype == torch.uint8 and storage.numel() >= self.nbytes and storage.data_ptr() % 256 == 0
        self.storage = storage
        self.codes = storage[offs[0]:offs[0] + sizes[0]].view(n, r3 // 4)
        self.lit_count = storage[offs[1]:offs[1] + sizes[1]].view(torch.int32)
        self.lits = storage[offs[2]:offs[2] + sizes[2]].view(torch.float32).view(n, 3 * r3)

    def rows(self, f0, f1):
        """The packed form of volumes [f0, f1) as an object of its own (a view: same memory)."""
        v = object.__new__(PackedVolumes)
        v.n_frames, v.voxel_res, v.nbytes, v.storage = int(f1) - int(f0), self.voxel_res, self.nbytes, self.storage
        v.codes, v.lit_count, v.lits = self.codes[f0:f1], self.lit_count[f0:f1], self.lits[f0:f1]
        return v

    @staticmethod
    def storage_bytes(n_frames, voxel_res=32):
        r3 = int(voxel_res) ** 3
        return sum(((sz + 255) & ~255) for sz in (n_frames * (r3 // 4), n_frames * 4, n_frames * 3 * r3 * 4)) or 256

    def bytes_in_use(self, n_frames=None):
        """Bytes a reader of the first ``n_frames`` packed volumes touches (synchronises)."""
        n = self.n_frames if n_frames is None else int(n_frames)
        r3 = self.voxel_res ** 3
        return n * (r3 // 4 + 4) + 12 * int(self.lit_count[:n].sum(dtype=torch.int64).item())


def pack_volumes_device(vol, packed=None, stream=None):
    """Dense float32[n,3,R,R,R] (either layout) -> ``PackedVolumes``, asynchronous on the current stream."""
    if not vol.is_cuda:
        raise RuntimeError("pack_volumes_device needs a CUDA tensor; there is no CPU fallback")
    assert vol.dtype == torch.float32 and vol.is_contiguous() and vol.dim() == 5 and vol.shape[1] == 3
    n, R = int(vol.shape[0]), int(vol.shape[2])
    if packed is None:
        packed = PackedVolumes(n, R, vol.device)
    assert packed.voxel_res == R and packed.n_frames >= n
    with torch.cuda.device(vol.device):
        check(load_library().dtsdf_pack_volumes_batch(R, n, vol.data_ptr(), packed.codes.data_ptr(),
                                                      packed.lit_count.data_ptr(), packed.lits.data_ptr(),
                                                      _stream_ptr(stream)))
    return packed


def unpack_volumes_device(packed, n_frames=None, out=None, device=None, stream=None):
    """``PackedVolumes`` (local or another GPU's, peer-mapped) -> dense float32[n,3,R,R,R] on ``device``
    (default: where ``out`` or the packed arrays live), asynchronous on the current stream."""
    n = packed.n_frames if n_frames is None else int(n_frames)
    R = packed.voxel_res
    if out is None:
        out = torch.empty((n, 3, R, R, R), dtype=torch.float32, device=device or packed.codes.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == n * 3 * R ** 3
    with torch.cuda.device(out.device):
        check(load_library().dtsdf_unpack_volumes_batch(R, n, packed.codes.data_ptr(), packed.lit_count.data_ptr(),
                                                        packed.lits.data_ptr(), out.data_ptr(), _stream_ptr(stream)))
    return out
