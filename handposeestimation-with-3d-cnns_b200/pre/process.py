"""Drop-in for the reference's ``pre/process.py``: same class, methods and return values,
computed by the sm_100a CUDA library (include/dtsdf_b200.h).  No CPU fallback.

Reference behaviour kept (paths relative to the reference root):
  DataProcess(data, ground_truth, point_num=6000, aug=False) ....... pre/process.py:6-11
  .process() -> [point_clouds, tsdf, max_l, mid_p] .................. pre/process.py:13-28
  .point_cloud(), .set_length(), .tsdf_f(), .max_min_point(), .tsdf_cal()  pre/process.py:30-200
  cut_version aliases: DataProcess(data, point_num=6000, aug=False), .get_points(),
  .normalize(gt, max_l, mid_p) ..................................... cut_version/pre/process.py:6,52,251

What runs where: back-projection + compaction, the gather by the resampling index vector,
the per-axis extrema, the grid parameters and the voxel loop run on the GPU.  The index
vector itself is drawn on the host from the *global* ``np.random`` stream with exactly
the reference's calls (pre/process.py:75-80), so seeding ``np.random`` reproduces the
reference's sampled cloud bit for bit.

Differences, all deliberate:
  * the volume is returned as float32 ``[3,32,32,32]`` (c,x,y,z); the reference returns the
    same values in a float64 array (every value it stores is a float32 scalar);
  * ``process()`` with ``aug=True`` raises NotImplementedError: the original's ``data_aug(pc)``
    itself raises ``AxisError`` (pre/process.py:252, ``np.mean(..., axis=2)`` on a 2-D cloud), so
    there is no behaviour to reproduce.  The later fork's ``data_aug(joints, point_clouds)``
    (cut_version/pre/process.py:69-134) does run and is reproduced bit for bit.
"""
import numpy as np

from _bind import package

_pkg = package()


def draw_aug():
    """The five random parameters of the fork's ``data_aug`` from the global ``np.random`` stream, in its
    order (cut_version/pre/process.py:76-94): stretch x, stretch y ~ U[2/3, 3/2), then whole degrees about
    x and y from [-10, 10) and about z from [-30, 30)."""
    sx, sy = np.random.uniform(2 / 3, 3 / 2), np.random.uniform(2 / 3, 3 / 2)
    return sx, sy, np.random.randint(-10, 10), np.random.randint(-10, 10), np.random.randint(-30, 30)


def _plane_rotation(i, j, c, s):
    """3x3 identity with [[c, s], [-s, c]] in rows/columns (i, j)."""
    m = np.eye(3)
    m[i, i], m[i, j], m[j, i], m[j, j] = c, s, -s, c
    return m


def apply_aug(params, joints, point_clouds):
    """The arithmetic of the fork's ``data_aug`` (cut_version/pre/process.py:79-134) for drawn parameters:
    ``S = diag(sx, sy, 1)``; ``R = Rx . Ry . Rz`` with the x and y angles converted to radians and the z
    angle NOT (``np.cos(rot_z)`` of a number of degrees, :95-96); every row ``p`` becomes
    ``(p . S) . R + mean(p . S)`` - the "centre" that is added back is the mean of the row's own three
    stretched coordinates (``np.mean(.., axis=1)``, :110,125), and it was never subtracted (:116,129 rotate the
    un-centred array).  float64 throughout, the fork's np.dot calls on the fork's shapes."""
    sx, sy, rot_x, rot_y, rot_z = params
    S = np.diag(np.array([sx, sy, 1]))
    ax, ay = np.pi * rot_x / 180, np.pi * rot_y / 180
    R_x = _plane_rotation(1, 2, np.cos(ax), np.sin(ax))
    R_y = _plane_rotation(2, 0, np.cos(ay), np.sin(ay))
    R_z = _plane_rotation(0, 1, np.cos(rot_z), np.sin(rot_z))
    R = np.dot(np.dot(R_x, R_y), R_z)
    joints = np.asarray(joints)
    if len(joints.shape) != 2:
        joints = joints.reshape(21, 3)

    def moved(rows):
        stretched = np.dot(rows, S)
        return np.dot(stretched, R) + np.mean(stretched, axis=1)[:, None]
    return moved(point_clouds), moved(joints)


class DataProcess(object):

    def __init__(self, data, ground_truth=None, point_num=6000, aug=False):
        # cut_version signature DataProcess(data, point_num=6000, aug=False)
        # (cut_version/pre/process.py:6): the second positional argument is the point count
        if isinstance(ground_truth, (int, np.integer)) and not isinstance(ground_truth, bool):
            if isinstance(point_num, (bool, np.bool_)):
                aug = bool(point_num)
            point_num = int(ground_truth)
            ground_truth = None
        self.fFocal_msra = 241.42
        self.data = data
        self.ground_truth = ground_truth
        self.point_num = point_num
        self.aug = aug
        self._vox = _pkg.default_voxelizer()

    # ------------------------------------------------------------------ pre/process.py:13-28
    def process(self):
        hand_points = self.point_cloud()
        point_clouds = self.set_length(hand_points)
        tsdf, max_l, mid_p = self.tsdf_f(point_clouds)
        if self.aug:
            hand_aug, ground_truth_aug = self.data_aug(hand_points)
            point_clouds_aug = self.set_length(hand_aug)
            tsdf_aug, max_l_aug, mid_p_aug = self.tsdf_f(point_clouds_aug)
            return [point_clouds, tsdf, max_l, mid_p,
                    point_clouds_aug, tsdf_aug, max_l_aug, mid_p_aug, ground_truth_aug]
        return [point_clouds, tsdf, max_l, mid_p]

    # ------------------------------------------------------------------ pre/process.py:30-68
    def point_cloud(self):
        "get point cloud from depth informations (float64[n_valid, 3])"
        return self._vox.point_cloud(self.data)

    # ------------------------------------------------------------------ pre/process.py:70-84
    def set_length(self, hand_points):
        "set the point number"
        point_num = self.point_num
        point_shape = hand_points.shape[0]
        if point_shape < point_num:
            rand_idx = np.arange(0, point_num, 1, dtype=np.int32)
            rand_idx[point_shape:] = np.random.randint(0, point_shape,
                                                       size=point_num - point_shape)
        else:
            rand_idx = np.random.randint(0, point_shape, size=point_num)
        # the array handed in is what gets gathered, like the reference's hand_points[rand_idx, :]: the
        # device context is shared by every DataProcess / tsdf_for call of the process, so nothing is
        # assumed to be resident from an earlier call (and in-place edits of hand_points count)
        pc, _, _ = self._vox.resample_extrema(point_num, rand_idx, points=hand_points)
        return pc

    # cut_version/pre/process.py:52-67
    def get_points(self):
        return self.set_length(self.point_cloud())

    # ------------------------------------------------------------------ pre/process.py:102-120
    def max_min_point(self, point_cloud):
        pc = np.asarray(point_cloud, dtype=np.float64)
        _, pmax, pmin = self._vox.resample_extrema(pc.shape[0], None, points=pc, want_points=False)
        return pmax, pmin

    # ------------------------------------------------------------------ pre/process.py:86-100
    def tsdf_f(self, point_cloud):
        point_max, point_min = self.max_min_point(point_cloud)
        aux = np.concatenate([point_max, point_min]).astype(np.float32).reshape(1, 6)
        vol, max_l, mid_p = self._vox.voxelize_frames(
            [self.data], layout=_pkg.LAYOUT_CXYZ, bbox_mode=_pkg.BBOX_EXTREMA, aux=aux)
        return vol[0], max_l[0], mid_p[0]

    # ------------------------------------------------------------------ pre/process.py:122-200
    def tsdf_cal(self, vox_ori, voxel_len, truncation):
        aux = np.empty((1, 5), dtype=np.float32)
        aux[0, :3] = np.asarray(vox_ori, dtype=np.float32)
        aux[0, 3] = np.float32(voxel_len)
        aux[0, 4] = np.float32(truncation)
        vol, _, _ = self._vox.voxelize_frames(
            [self.data], layout=_pkg.LAYOUT_CXYZ, bbox_mode=_pkg.BBOX_GRID, aux=aux)
        return vol[0]

    # ------------------------------------------------------------------ pre/process.py:202-261,
    #                                                                    cut_version/pre/process.py:69-134
    def data_aug(self, *args):
        """``data_aug(joints, point_clouds)`` - the later fork's augmentation, the one variant of the
        reference that runs (cut_version/pre/process.py:69-134): returns ``(point_clouds_aug, ground_truth_aug)``
        in float64 with the fork's own NumPy calls in the fork's order, quirks included (see ``draw_aug`` /
        ``apply_aug``), so a seeded run reproduces it bit for bit.  It is host float64 linear algebra on a
        6000 x 3 cloud; the D-TSDF of the augmented cloud is then ``tsdf_f(point_clouds_aug)`` - the fused
        kernel with the bounding box of the augmented points over the UN-augmented depth map, as the fork
        does (its own note at :207).

        ``data_aug(point_clouds)`` - the original's signature (pre/process.py:202): that implementation
        raises ``AxisError`` (:252, ``np.mean(..., axis=2)`` of a 2-D array) before returning anything, so
        there is no behaviour to reproduce: NotImplementedError."""
        if len(args) == 2:
            joints, point_clouds = args
            return apply_aug(draw_aug(), joints, point_clouds)
        raise NotImplementedError(
            "data_aug(point_clouds): the reference implementation raises AxisError (pre/process.py:252); "
            "the fork's data_aug(joints, point_clouds) (cut_version/pre/process.py:69) is implemented, and a "
            "consistent on-device augmentation is specified separately (SURVEY.md section 8f, N2)")

    # ------------------------------------------------------------------ cut_version/pre/process.py:251-263
    def normalize(self, ground_truth, max_l, mid_p):
        gt = np.asarray(ground_truth, dtype=np.float32).reshape(-1, 63)
        n = gt.shape[0]
        ml = np.broadcast_to(np.asarray(max_l, dtype=np.float32).reshape(-1), (n,))
        mp = np.broadcast_to(np.asarray(mid_p, dtype=np.float32).reshape(-1, 3), (n, 3))
        return self._vox.normalize_joints(gt, ml, mp)
