"""Volumes of every rank in ONE buffer in the root rank's HBM, filled by the voxeliser kernels
themselves (the optional gather of volumes into a training rank).

The reference has no multi-GPU code at all (single process, pre/read_MSRA.py:98-111 loops over
frames one by one).  Here frames shard across ranks with no exchange while voxelising; what a
training rank wants afterwards is all volumes in its own memory.  Instead of "kernel writes local
HBM, then a collective copies it" (write + read + NVLink + write), the root rank shares its output
buffer over CUDA IPC (``dtsdf_shared_alloc`` / ``dtsdf_shared_open`` of the C ABI) and every other
rank hands the fused kernel a pointer INTO that buffer: the kernel's 128-byte line stores travel
over NVLink 5 / NVSwitch as posted peer writes and land in the root's HBM exactly once.  The only
collective left is a barrier.

torch is used for what it is here for: wrapping device memory and the process group.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .voxelizer import shard_range


def enable_peer_access(dev, peer):
    """Lets kernels running on `dev` address memory of `peer` (both device indices of this process):
    ``dtsdf_enable_peer_access`` of the C ABI."""
    capi.check(capi.load_library().dtsdf_enable_peer_access(int(dev), int(peer)))


class _DevicePointer:
    """Raw device memory as a ``__cuda_array_interface__`` object (for ``torch.as_tensor``)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerVolumeBuffer:
    """``full`` = float32[n_total,3,R,R,R] in the root rank's memory, mapped in every rank;
    ``mine`` = this rank's frame range of it (``shard_range``), to be passed as ``out_vol`` to
    ``voxelize_device``.  After the kernels, ``complete()`` makes the root's view consistent.
    ``close()`` must be called on every rank (a collective: the owner frees last)."""

    def __init__(self, n_total, voxel_res=32, root=0, group=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.root = root
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        R = int(voxel_res)
        shape = (int(n_total), 3, R, R, R)
        nbytes = int(np.prod(shape)) * 4
        lib = capi.load_library()
        self._ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        if self.rank == root:
            capi.check(lib.dtsdf_shared_alloc(self.device.index, nbytes, ctypes.byref(self._ptr), handle))
            payload = [bytes(handle)]
        else:
            payload = [None]
        dist.broadcast_object_list(payload, src=root, group=group)
        if self.rank != root:
            handle = (ctypes.c_uint8 * 64).from_buffer_copy(payload[0])
            capi.check(lib.dtsdf_shared_open(self.device.index, handle, ctypes.byref(self._ptr)))
        with torch.cuda.device(self.device):
            self.full = torch.as_tensor(_DevicePointer(self._ptr.value, shape), device=self.device)
        lo, hi = shard_range(n_total, self.rank, self.world)
        self.lo, self.hi = lo, hi
        self.mine = self.full[lo:hi]

    def complete(self):
        """Every rank's kernels are done and their stores visible to the root."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def close(self):
        if self._ptr is None:
            return
        self.mine = self.full = None
        torch.cuda.synchronize(self.device)
        lib = capi.load_library()
        if self.rank != self.root:
            capi.check(lib.dtsdf_shared_close(self.device.index, self._ptr))
        dist.barrier(group=self.group)
        if self.rank == self.root:
            capi.check(lib.dtsdf_shared_free(self.device.index, self._ptr))
        self._ptr = None
