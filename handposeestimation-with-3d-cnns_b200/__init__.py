"""B200-native projective directional TSDF (D-TSDF) voxeliser.

Drop-in for the depth-crop -> 3 x R^3 volume path of
Moon0shang/HandPoseEstimation-with-3D-CNNs (pre/process.py, pre/tsdf_for.py,
pre/tsdf_numba.py and the 3D_CNN/dataset.py consumer).  The directory name is not a
Python identifier; load it with ``__graft_entry__.load_package()`` (alias
``hpe3d_b200``) or, like the reference's own scripts, put ``pre/`` on ``sys.path`` and
``from process import DataProcess``.
"""
from . import capi, synth
from .capi import (BBOX_EXTREMA, BBOX_GRID, BBOX_PIXELS, LAYOUT_CXYZ, LAYOUT_CZYX, DtsdfError)
from .voxelizer import (PackedHostVolumeList, PackedHostVolumes, Voxelizer, default_voxelizer, normalize_joints_device, pack_frames,
                        shard_packed, shard_range, voxelize_device, voxelize_sampled_device, count_valid_device)

__all__ = [
    "capi", "synth", "Voxelizer", "PackedHostVolumes", "PackedHostVolumeList", "default_voxelizer", "pack_frames", "voxelize_device",
    "normalize_joints_device", "voxelize_sampled_device", "count_valid_device", "shard_range", "shard_packed", "DtsdfError",
    "LAYOUT_CXYZ", "LAYOUT_CZYX", "BBOX_PIXELS", "BBOX_EXTREMA", "BBOX_GRID",
]
