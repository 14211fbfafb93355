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
from .codec import PackedVolumes, pack_volumes_device, unpack_volumes_device
from .voxelizer import shard_range


def enable_peer_access(dev, peer):
    """Lets kernels running on `dev` address memory of `peer` (both device indices of this process):
    ``dtsdf_enable_peer_access`` of the C ABI."""
    capi.check(capi.load_library().dtsdf_enable_peer_access(int(dev), int(peer)))


class _DevicePointer:
    """Raw device memory as a ``__cuda_array_interface__`` object (for ``torch.as_tensor``)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
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


class PackedVolumeGather:
    """The same gather with an eighth of the bytes on NVLink: every rank packs its volumes
    (``codec.pack_volumes_device``: 2-bit codes + literal triples, lossless) into a buffer of its OWN
    memory that the root has mapped (CUDA IPC), and the root's expansion kernel
    (``dtsdf_unpack_volumes_batch``, one CTA per frame) pulls each frame's ~47 KB in use straight out of
    the sender's HBM with TMA bulk copies while it writes the dense 393 KB into ``full`` - transfer and
    expansion are one kernel on the receiving rank, and nothing but two stream-ordered barriers
    (one-element all-reduces) is left of the collective.

    What bounds it: the root writes every volume of every rank into its own HBM (n_total x 393 KB; at 8 x
    8,500 frames 26.7 GB = 4.2 ms at the measured copy peak), so a gather into one GPU can at best run at
    about one GPU's own voxelising rate; the dense gather is bound by the root's NVLink ingress instead
    (23.4 GB at ~0.7 TB/s = 33 ms).

    ``frames_per_rank[r]`` volumes come from rank r; ``full`` (root only) is float32[sum,3,R,R,R] and the
    root writes its own range of it itself (``mine``: pass it as ``out_vol`` to ``voxelize_device``).
    ``close()`` on every rank (a collective)."""

    def __init__(self, frames_per_rank, voxel_res=32, root=0, group=None, device=None, full=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.root = int(root)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.counts = [int(c) for c in frames_per_rank]
        assert len(self.counts) == self.world
        self.R = int(voxel_res)
        self.starts = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)
        lib = capi.load_library()
        self._lib = lib
        n_mine = self.counts[self.rank]
        self._ptr = ctypes.c_void_p()
        self._peer_ptrs = {}
        self.packed = None
        handle = (ctypes.c_uint8 * 64)()
        payload = None
        if self.rank != self.root and n_mine > 0:
            nbytes = PackedVolumes.storage_bytes(n_mine, self.R)
            capi.check(lib.dtsdf_shared_alloc(self.device.index, nbytes, ctypes.byref(self._ptr), handle))
            payload = bytes(handle)
            self.packed = PackedVolumes(n_mine, self.R, self.device, storage=self._as_bytes(self._ptr.value, nbytes))
        handles = [None] * self.world
        dist.all_gather_object(handles, payload, group=group)
        self.peers = {}
        self.full = self.mine = None
        if self.rank == self.root:
            for r, h in enumerate(handles):
                if h is None:
                    continue
                p = ctypes.c_void_p()
                hb = (ctypes.c_uint8 * 64).from_buffer_copy(h)
                capi.check(lib.dtsdf_shared_open(self.device.index, hb, ctypes.byref(p)))
                self._peer_ptrs[r] = p
                nbytes = PackedVolumes.storage_bytes(self.counts[r], self.R)
                self.peers[r] = PackedVolumes(self.counts[r], self.R, self.device,
                                              storage=self._as_bytes(p.value, nbytes))
            n_total = int(self.starts[-1])
            self.full = full if full is not None else torch.empty((n_total, 3, self.R, self.R, self.R),
                                                                  dtype=torch.float32, device=self.device)
            self.mine = self.full[int(self.starts[self.rank]):int(self.starts[self.rank + 1])]
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _as_bytes(self, ptr, nbytes):
        with torch.cuda.device(self.device):
            return torch.as_tensor(_DevicePointer(ptr, (int(nbytes),), "|u1"), device=self.device)

    def _stream_barrier(self):
        """Orders the work enqueued so far on every rank's current stream before what follows on every
        rank's current stream; the host does not wait."""
        dist.all_reduce(self._flag, group=self.group)

    def pack(self, vol):
        """Non-root ranks: pack ``vol`` (float32[n_mine,3,R,R,R]) into the shared buffer (current stream)."""
        if self.rank != self.root and self.packed is not None:
            pack_volumes_device(vol, self.packed)

    def gather(self):
        """Collective.  Before: every non-root rank has enqueued ``pack`` and the root has enqueued the
        writes of its own range ``mine``.  After (on the root's current stream): ``full`` is complete.
        The packed buffers may be rewritten once this call has been enqueued on every rank."""
        self._stream_barrier()
        if self.rank == self.root:
            for r, pk in self.peers.items():
                unpack_volumes_device(pk, out=self.full[int(self.starts[r]):int(self.starts[r + 1])])
        self._stream_barrier()
        return self.full

    # ---- the same in parts: the root expands part k of every rank while the others produce part k + 1
    def part(self, r, k, parts):
        """Frames [f0, f1) of rank r that make up part k of ``parts``."""
        n = self.counts[r]
        return n * k // parts, n * (k + 1) // parts

    def gather_part(self, k, parts, vol=None):
        """Collective, once per part and in order.  Non-root ranks: ``vol`` holds the rank's dense volumes with part
        k written (current stream); it is packed into the same rows of the shared buffer.  Root: its own part k is
        enqueued already; after the barrier it expands part k of every other rank.  With two parts the root's
        expansion of the first half covers the other ranks' voxelising and packing of the second - what is left
        of the step is the root's own voxelisation plus its expansion of everything.  ``finish_parts()`` after the
        last part."""
        if self.rank != self.root and self.packed is not None:
            f0, f1 = self.part(self.rank, k, parts)
            if f1 > f0:
                pack_volumes_device(vol[f0:f1], self.packed.rows(f0, f1))
        self._stream_barrier()
        if self.rank == self.root:
            for r, pk in self.peers.items():
                f0, f1 = self.part(r, k, parts)
                if f1 > f0:
                    s = int(self.starts[r])
                    unpack_volumes_device(pk.rows(f0, f1), out=self.full[s + f0:s + f1])

    def finish_parts(self):
        """Collective: the packed buffers may be rewritten once this has been enqueued on every rank."""
        self._stream_barrier()
        return self.full

    def bytes_on_nvlink(self):
        """Bytes the root pulls from the other ranks per gather (a collective; synchronises)."""
        mine = self.packed.bytes_in_use() if self.packed is not None else 0
        t = torch.tensor([float(mine)], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self):
        if self._lib is None:
            return
        torch.cuda.synchronize(self.device)
        self.full = self.mine = None
        self.peers = {}
        for p in self._peer_ptrs.values():
            capi.check(self._lib.dtsdf_shared_close(self.device.index, p))
        self._peer_ptrs = {}
        dist.barrier(group=self.group)
        self.packed = None
        if self._ptr.value:
            capi.check(self._lib.dtsdf_shared_free(self.device.index, self._ptr))
        self._lib = None
