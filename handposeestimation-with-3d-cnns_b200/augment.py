"""Training-time 3-D augmentation of the D-TSDF.

The reference draws a stretch ``S = diag(sx, sy, 1)``, ``sx, sy ~ U[2/3, 3/2]`` and a rotation ``R``
and forms ``pc . S . R`` (show/visualize.py:59-119; pre/process.py:202-261 is the same idea but raises
AxisError at :252, and cut_version/pre/process.py:69-134 then voxelises the *un-augmented* depth map,
:207).  Here the augmentation is applied consistently: the voxel grid is laid over the augmented
hand and every voxel looks its depth up through the inverse map (``dtsdf_voxelize_aug_batch``,
specified by oracle/dtsdf_oracle.c:oracle_dtsdf_aug_frame).
"""
import ctypes

import numpy as np

from . import capi
from .capi import check, load_library


def make_augmentations(n, rng=None, stretch=(2.0 / 3.0, 3.0 / 2.0), max_rot_xy_deg=30.0, max_rot_z_deg=180.0,
                       centers=None):
    """``float32[n,21]`` = A (3x3), Ainv (3x3), c (3) per frame, with the reference's parameter ranges
    (pre/process.py:204-224 docstring: stretch x,y in [2/3,3/2], z 1; rotation about x,y in [-30,30]
    degrees, about z in [-180,180]) drawn from ``rng`` (a ``np.random.Generator``; default seed 0).
    ``A = S . Rx . Ry . Rz`` for row vectors ``p' = (p - c) . A + c``; ``centers`` float[n,3] (default 0,
    as show/visualize.py).  ``Ainv`` is the inverse of the float32-rounded ``A``."""
    rng = np.random.default_rng(0) if rng is None else rng
    out = np.zeros((n, 21), dtype=np.float32)
    # five uniform draws per frame, frame after frame (sx, sy, ax, ay, az), whole batch at once
    u = rng.random((n, 5))
    sx = stretch[0] + (stretch[1] - stretch[0]) * u[:, 0]
    sy = stretch[0] + (stretch[1] - stretch[0]) * u[:, 1]
    ax = np.deg2rad(-max_rot_xy_deg + 2.0 * max_rot_xy_deg * u[:, 2])
    ay = np.deg2rad(-max_rot_xy_deg + 2.0 * max_rot_xy_deg * u[:, 3])
    az = np.deg2rad(-max_rot_z_deg + 2.0 * max_rot_z_deg * u[:, 4])
    S = np.zeros((n, 3, 3)); Rx = np.zeros((n, 3, 3)); Ry = np.zeros((n, 3, 3)); Rz = np.zeros((n, 3, 3))
    S[:, 0, 0], S[:, 1, 1], S[:, 2, 2] = sx, sy, 1.0
    Rx[:, 0, 0] = 1.0; Rx[:, 1, 1] = np.cos(ax); Rx[:, 1, 2] = np.sin(ax); Rx[:, 2, 1] = -np.sin(ax); Rx[:, 2, 2] = np.cos(ax)
    Ry[:, 1, 1] = 1.0; Ry[:, 0, 0] = np.cos(ay); Ry[:, 0, 2] = -np.sin(ay); Ry[:, 2, 0] = np.sin(ay); Ry[:, 2, 2] = np.cos(ay)
    Rz[:, 2, 2] = 1.0; Rz[:, 0, 0] = np.cos(az); Rz[:, 0, 1] = np.sin(az); Rz[:, 1, 0] = -np.sin(az); Rz[:, 1, 1] = np.cos(az)
    A = (S @ Rx @ Ry @ Rz).astype(np.float32)
    out[:, :9] = A.reshape(n, 9)
    if n:
        out[:, 9:18] = np.linalg.inv(A.astype(np.float64)).astype(np.float32).reshape(n, 9)
    if centers is not None:
        out[:, 18:] = np.asarray(centers, dtype=np.float32).reshape(n, 3)
    return out


def identity_augmentations(n):
    out = np.zeros((n, 21), dtype=np.float32)
    out[:, [0, 4, 8, 9, 13, 17]] = 1.0
    return out


def voxelize_aug_device(depth_packed, offsets, headers, aug, max_crop_px=0, cfg=None, out_vol=None, out_max_l=None,
                        out_mid_p=None, stream=None, joints=None, joint_flags=0, out_joints=None):
    """``dtsdf_voxelize_aug_batch`` on torch CUDA tensors (asynchronous on torch's current stream).
    ``aug`` float32[n,21] on the device.  Returns (vol, max_l, mid_p[, joints_aug_normalised])."""
    import torch
    from .voxelizer import _torch_stream_ptr
    if not (depth_packed.is_cuda and offsets.is_cuda and headers.is_cuda and aug.is_cuda):
        raise RuntimeError("voxelize_aug_device needs CUDA tensors; there is no CPU fallback")
    assert aug.dtype == torch.float32 and aug.is_contiguous()
    cfg = cfg or capi.default_config()
    n, R, dev = int(headers.shape[0]), int(cfg.voxel_res), depth_packed.device
    assert aug.numel() == 21 * n
    if out_vol is None:
        out_vol = torch.empty((n, 3, R, R, R), dtype=torch.float32, device=dev)
    if out_max_l is None:
        out_max_l = torch.empty((n,), dtype=torch.float32, device=dev)
    if out_mid_p is None:
        out_mid_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    jp = op = None
    if joints is not None:
        assert joints.is_cuda and joints.dtype == torch.float32 and joints.is_contiguous() and joints.numel() == 63 * n
        if out_joints is None:
            out_joints = torch.empty((n, 63), dtype=torch.float32, device=dev)
        jp, op = joints.data_ptr(), out_joints.data_ptr()
    with torch.cuda.device(dev):
        check(load_library().dtsdf_voxelize_aug_batch(
            ctypes.byref(cfg), n, depth_packed.data_ptr(), offsets.data_ptr(), headers.data_ptr(), aug.data_ptr(),
            int(max_crop_px), jp, int(joint_flags), out_vol.data_ptr(), out_max_l.data_ptr(), out_mid_p.data_ptr(),
            op, _torch_stream_ptr(stream)))
    if joints is not None:
        return out_vol, out_max_l, out_mid_p, out_joints
    return out_vol, out_max_l, out_mid_p
