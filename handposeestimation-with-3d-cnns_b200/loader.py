"""On-the-fly D-TSDF batches for the 3D CNN, voxelised on a side stream while the consumer works
(BASELINE.json configs[4]).

The reference trains from volumes that were preprocessed to disk, loaded whole into host RAM
(30 GB for the full MSRA set, 3D_CNN/dataset.py:35,99-117) and copied to the GPU batch by batch
(3D_CNN/train.py:232), and normalises the joints in a Python loop per sample (:236-244).  On a
180 GB B200 the *inputs* of the voxeliser fit in HBM outright (76k crops = 4.8 GB), so:

  * ``ResidentFrames`` keeps the packed depth crops, headers, offsets and joints of the whole
    data set in device memory; a batch is an index gather of two small tables, no crop moves;
  * ``OnTheFlyLoader`` iterates (optionally shuffled) batches: batch k+1 is voxelised by the fused
    kernel - joint normalisation in the same launch - on a side stream into the second of two
    output buffers while the consumer's work on batch k runs on the current stream.
"""
import numpy as np
import torch

from . import capi
from .augment import make_augmentations, voxelize_aug_device
from .voxelizer import pack_frames, voxelize_device


class ResidentFrames:
    """Packed frames (include/dtsdf_b200.h, "Packed batch") resident on one GPU."""

    def __init__(self, headers, offsets, packed, joints=None, device="cuda", max_crop_px=None):
        dev = torch.device(device)
        self.device = dev
        self.headers = torch.as_tensor(np.ascontiguousarray(headers, dtype=np.int32)).to(dev)
        self.offsets = torch.as_tensor(np.ascontiguousarray(offsets, dtype=np.int64)).to(dev)
        self.packed = torch.as_tensor(np.ascontiguousarray(packed, dtype=np.float32)).to(dev)
        self.joints = None
        if joints is not None:
            self.joints = torch.as_tensor(np.ascontiguousarray(joints, dtype=np.float32).reshape(-1, 63)).to(dev)
        if max_crop_px is None:
            h = np.asarray(headers).reshape(-1, 6)
            max_crop_px = int(np.max((h[:, 4] - h[:, 2]).astype(np.int64) * (h[:, 5] - h[:, 3]))) if len(h) else 0
        self.max_crop_px = int(max_crop_px)

    @classmethod
    def from_frames(cls, frames, joints=None, device="cuda"):
        headers, offsets, packed, max_px = pack_frames(frames)
        return cls(headers, offsets, packed, joints, device, max_px)

    def __len__(self):
        return int(self.headers.shape[0])

    def voxelize(self, index, cfg=None, joint_flags=capi.JOINT_CLAMP01, out=None, stream=None, aug=None):
        """Volumes (and normalised joints) of the frames ``index`` (int64 tensor on the device), on
        ``stream`` (default: the current one).  ``out`` = preallocated (vol, max_l, mid_p[, joints]).
        ``aug`` float32[len(index),21] on the device: augmented volumes and joints (augment.py)."""
        hdr = self.headers.index_select(0, index)
        off = self.offsets.index_select(0, index)
        o = out or (None, None, None, None)
        if aug is not None:
            j = self.joints.index_select(0, index) if self.joints is not None else None
            return voxelize_aug_device(self.packed, off, hdr, aug, self.max_crop_px, cfg, out_vol=o[0], out_max_l=o[1],
                                       out_mid_p=o[2], stream=stream, joints=j, joint_flags=joint_flags,
                                       out_joints=o[3] if (j is not None and len(o) > 3) else None)
        if self.joints is not None:
            return voxelize_device(self.packed, off, hdr, self.max_crop_px, cfg, out_vol=o[0], out_max_l=o[1],
                                   out_mid_p=o[2], stream=stream, joints=self.joints.index_select(0, index),
                                   joint_flags=joint_flags, out_joints=o[3] if len(o) > 3 else None)
        return voxelize_device(self.packed, off, hdr, self.max_crop_px, cfg, out_vol=o[0], out_max_l=o[1],
                               out_mid_p=o[2], stream=stream)


class OnTheFlyLoader:
    """``for tsdf, joints_nor, max_l, mid_p in loader:`` - three output buffers: the tensors of batch k
    stay valid (and may be used by work enqueued on the consumer's current stream) during iteration k
    AND iteration k+1, i.e. until batch k+2 is requested; batch k+1 is voxelised on a side stream while
    the consumer works on batch k.  Ordering is by CUDA events on whatever stream is current where the
    consumer asks for a batch: that stream waits for the batch at yield time, and the buffer of batch
    k-1 is released on it when the consumer comes back for batch k+1.  ``joints_nor`` is None without
    ground truth."""

    NBUF = 3

    def __init__(self, frames, batch_size=16, shuffle=False, seed=0, voxel_res=32, layout=capi.LAYOUT_CXYZ,
                 joint_flags=capi.JOINT_CLAMP01, drop_last=False, augment=False, augment_center=None):
        """``augment=True``: every batch gets fresh random stretch + rotation (augment.make_augmentations,
        the reference's parameter ranges) about ``augment_center`` (3 floats, default the origin)."""
        self.frames = frames
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.seed = int(seed)
        self.epoch = 0
        self.cfg = capi.default_config(voxel_res, layout, capi.BBOX_PIXELS)
        self.joint_flags = int(joint_flags)
        self.drop_last = drop_last
        self.augment = bool(augment)
        self.augment_center = None if augment_center is None else np.asarray(augment_center, np.float32).reshape(1, 3)
        self._aug_rng = np.random.default_rng(self.seed + 7919)
        self.last_aug = [None] * self.NBUF
        R, B, dev = int(voxel_res), self.batch_size, frames.device
        self._bufs = []
        for _ in range(self.NBUF):
            self._bufs.append((torch.empty((B, 3, R, R, R), dtype=torch.float32, device=dev),
                               torch.empty((B,), dtype=torch.float32, device=dev),
                               torch.empty((B, 3), dtype=torch.float32, device=dev),
                               torch.empty((B, 63), dtype=torch.float32, device=dev)))
        self._side = torch.cuda.Stream(device=dev)

    def __len__(self):
        n = len(self.frames)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _order(self):
        n = len(self.frames)
        if self.shuffle:
            g = torch.Generator(device="cpu")
            g.manual_seed(self.seed + self.epoch)
            return torch.randperm(n, generator=g).to(self.frames.device)
        return torch.arange(n, device=self.frames.device)

    def __iter__(self):
        fr, B, side = self.frames, self.batch_size, self._side
        order = self._order()
        self.epoch += 1
        nb = len(self)
        NB = self.NBUF
        ready = [torch.cuda.Event() for _ in range(NB)]
        consumed = [torch.cuda.Event() for _ in range(NB)]
        start = torch.cuda.current_stream(fr.device)
        side.wait_stream(start)                      # `order` and the resident tables are ready
        for e in consumed:
            e.record(start)

        def launch(k):
            slot = k % NB
            idx = order[k * B:(k + 1) * B]
            m = int(idx.numel())
            aug = None
            if self.augment:
                centers = None if self.augment_center is None else np.repeat(self.augment_center, m, axis=0)
                aug_h = make_augmentations(m, self._aug_rng, centers=centers)
                self.last_aug[slot] = aug_h
            with torch.cuda.stream(side):
                side.wait_event(consumed[slot])      # the consumer is done with this buffer
                if self.augment:
                    aug = torch.from_numpy(aug_h).to(fr.device, non_blocking=True)
                buf = self._bufs[slot]
                out = tuple(t[:m] for t in buf)
                fr.voxelize(idx, self.cfg, self.joint_flags, out=out, stream=side, aug=aug)
                ready[slot].record(side)
            return m

        sizes = {0: launch(0)} if nb else {}
        for k in range(nb):
            if k + 1 < nb:
                sizes[k + 1] = launch(k + 1)         # overlaps the consumer's work on batch k
            slot = k % NB
            torch.cuda.current_stream(fr.device).wait_event(ready[slot])
            m = sizes.pop(k)
            vol, max_l, mid_p, jn = (t[:m] for t in self._bufs[slot])
            yield vol, (jn if fr.joints is not None else None), max_l, mid_p
            # the consumer is back for batch k+1: whatever it enqueued on batch k-1 (during iterations
            # k-1 and k) precedes this point of its stream; batch k+2 will reuse that buffer
            if k >= 1:
                consumed[(k - 1) % NB].record(torch.cuda.current_stream(fr.device))
