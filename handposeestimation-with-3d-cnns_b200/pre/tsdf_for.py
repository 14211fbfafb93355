"""Drop-in for the reference's ``pre/tsdf_for.py`` (free-function copy of the voxel
loop), computed by the sm_100a CUDA library.  No CPU fallback.

  tsdf_f(data, point_cloud) -> (tsdf_v, max_lenth, mid_point) .... pre/tsdf_for.py:6-20
  max_min_point(point_cloud) -> (point_max, point_min) ........... pre/tsdf_for.py:23-41
  tsdf_cal(data, vox_ori, voxel_len, truncation) -> tsdf_v ....... pre/tsdf_for.py:44-122

``data`` is ``{'header': int32[6], 'depth': float32[bw*bh]}``; the volume comes back as
float32 ``[3,32,32,32]`` indexed (c,x,y,z) with the values the reference stores in its
float64 array.
"""
import numpy as np

from _bind import package

_pkg = package()

fFocal_msra = 241.42


def max_min_point(point_cloud):
    pc = np.asarray(point_cloud, dtype=np.float64)
    _, pmax, pmin = _pkg.default_voxelizer().resample_extrema(
        pc.shape[0], None, points=pc, want_points=False)
    return pmax, pmin


def tsdf_f(data, point_cloud):
    point_max, point_min = max_min_point(point_cloud)
    aux = np.concatenate([point_max, point_min]).astype(np.float32).reshape(1, 6)
    vol, max_l, mid_p = _pkg.default_voxelizer().voxelize_frames(
        [data], layout=_pkg.LAYOUT_CXYZ, bbox_mode=_pkg.BBOX_EXTREMA, aux=aux)
    return vol[0], max_l[0], mid_p[0]


def tsdf_cal(data, vox_ori, voxel_len, truncation):
    aux = np.empty((1, 5), dtype=np.float32)
    aux[0, :3] = np.asarray(vox_ori, dtype=np.float32)
    aux[0, 3] = np.float32(voxel_len)
    aux[0, 4] = np.float32(truncation)
    vol, _, _ = _pkg.default_voxelizer().voxelize_frames(
        [data], layout=_pkg.LAYOUT_CXYZ, bbox_mode=_pkg.BBOX_GRID, aux=aux)
    return vol[0]
