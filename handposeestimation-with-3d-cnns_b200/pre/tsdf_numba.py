"""Drop-in for the reference's ``pre/tsdf_numba.py``: ``cal_tsdf_cuda(s)`` with the same
arguments, return value and error convention, on the sm_100a CUDA library instead of
numba.cuda.  No CPU fallback.

  cal_tsdf_cuda(s) ................................................. pre/tsdf_numba.py:119-171
    s = {'header': int32[6], 'data': float32[bw*bh]}  (key 'data', :132-133)
    -> (tsdf float32[3,R,R,R] indexed (c,z,y,x) (:70-72), max_l, mid_p)
    -> None after printing diagnostics when the bounding box degenerates (the reference
       returns None from its ``except RuntimeWarning`` branch, :162-171)
  module constants FOCAL, CENTER_X, CENTER_Y (:8-10), VOXEL_RES (the reference's missing
  ``params`` module, :3)

The reference needs two kernel launches, four PCIe transfers and a host NumPy reduction per
frame (:133-158); here the bounding box, the grid parameters and the voxel kernel are one
fused launch.  ``cal_tsdf_cuda_batch`` is the batched form of the same call.
"""
import numpy as np

from _bind import package

_pkg = package()

FOCAL = 241.42
CENTER_X = 160
CENTER_Y = 120
VOXEL_RES = 32
threadsperblock = 32
mm_threadsperblock = 128


def _voxelizer():
    return _pkg.default_voxelizer()


def cal_tsdf_cuda(s):
    frame = {"header": s["header"], "data": s["data"]}
    vol, max_l, mid_p = _voxelizer().voxelize_frames(
        [frame], layout=_pkg.LAYOUT_CZYX, bbox_mode=_pkg.BBOX_PIXELS, voxel_res=VOXEL_RES)
    max_l0, mid_p0 = max_l[0], mid_p[0]
    if not (np.isfinite(max_l0) and max_l0 > 0 and max_l0 < 99999 and np.all(np.isfinite(mid_p0))):
        # no valid pixel (or non-finite depth): the reference's host arithmetic warns here
        print("warning caught: ", "degenerate bounding box")
        print("mid:", mid_p0)
        print("max_l", max_l0)
        return None
    return vol[0], max_l0, mid_p0


def cal_tsdf_cuda_batch(frames, layout=None):
    """Batched ``cal_tsdf_cuda``: list of ``s`` dicts -> (tsdf [n,3,R,R,R], max_l [n], mid_p [n,3])."""
    layout = _pkg.LAYOUT_CZYX if layout is None else layout
    return _voxelizer().voxelize_frames(frames, layout=layout, bbox_mode=_pkg.BBOX_PIXELS,
                                        voxel_res=VOXEL_RES)
