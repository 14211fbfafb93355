"""Drop-in for the reference's ``pre/read_MSRA.py``: the offline preprocessing driver that turns
the MSRA hand-gesture data base into the ``result/`` tree the trainer loads.  Same module
constants, functions, on-disk formats and ``np.random`` consumption; the per-frame work runs on
the sm_100a CUDA library, one gesture (about 500 frames) per batch.  No CPU fallback.

  DB_dir, SAVE_dir, AUG ............................................. pre/read_MSRA.py:31-34
  main()   subjects x gestures x frames -> result/<sub>/{Point_Cloud,TSDF,ground_truth,num}/<ges>
           and result/data_num-<sub>.npy ............................ pre/read_MSRA.py:37-140
  read_joint(f_dir) -> (bin_num, ground_truth float32[bin_num,63]) ... pre/read_MSRA.py:143-152
  read_bin(f_name)  -> (header int32[6], depth float32[bw*bh]) ....... pre/read_MSRA.py:155-164

What changes: the reference calls ``DataProcess(...).process()`` once per ``.bin`` file (Python
loops over every pixel and every voxel, about 0.16 s per frame); here all frames of a gesture go
through ``Voxelizer.process_frames`` (dtsdf_process_begin_host / dtsdf_process_finish_host).  The
resampling index vectors are still drawn frame by frame from the global ``np.random`` stream, so
with the same seed the tree is bit-identical to the reference's.

``AUG = True``: with ``JOINT_NOR = False`` (the original's tree) it raises, as the original's ``data_aug`` does
(pre/process.py:252); with ``JOINT_NOR = True`` it writes the later fork's tree, ``_aug`` half included, bit
for bit (cut_version/pre/read_MSRA.py:46-124; tests/golden/make_golden_tree_cut.py runs the unmodified fork).

Not reproduced: the trailing ``joint_pca(aug=False)`` of the reference's ``__main__`` block (pre/joint_pca.py
is self-declared untested and passes a function where data is expected, :34,39).
"""
import os
import os.path

import numpy as np

from _bind import package

_pkg = package()

# files directions
DB_dir = '/home/x/DB/MSRA HandPoseDataset/cvpr15_MSRAHandGestureDB/'
SAVE_dir = './result'
# set whether to do data augmentation
AUG = False

# True: the tree of the later fork (cut_version/pre/read_MSRA.py:60-113): the first 3 subjects (:27), ground
# truth saved as [n,21,3] (:63), an extra joint_nor/<ges>.npy = (gt - mid_p)/max_l + 0.5 per frame in float64
# (cut_version/pre/process.py:251-263, 63 numbers per frame: NumPy on the host, so the file matches
# the fork's float64 arithmetic), no data_num-<sub>.npy (:131); with AUG = True also the _aug directories
# (:46-55,68-75,95-102,115-124)
JOINT_NOR = False

POINTS_NUM = 6000
_TREE = ('Point_Cloud', 'TSDF', 'ground_truth', 'num')


def _ensure_dir(path, label=None):
    """mkdir that reports like the reference does (pre/read_MSRA.py:39-59) but never hides other errors."""
    if os.path.isdir(path):
        if label:
            print('directory "%s" already exist!' % label)
        return
    os.makedirs(path)
    if label:
        print('create directory "%s".' % label)


def _load_gesture(folder):
    """All frame records of one gesture folder plus its ground truth."""
    count, joints = read_joint(folder)
    records = []
    for k in range(count):
        hdr, crop = read_bin(os.path.join(folder, '%06d_depth.bin' % k))
        records.append({'header': hdr, 'depth': crop})
    return count, joints, records


def _store_gesture(out_dir, ges, count, joints, clouds, volumes, sides, centres):
    """The four files of a gesture (pre/read_MSRA.py:119-126).  The reference accumulates into float64
    arrays (np.empty default dtype, :93-96), so that is what goes to disk."""
    f64 = np.float64
    np.save(os.path.join(out_dir, _TREE[0], ges + '.npy'), clouds)
    np.savez(os.path.join(out_dir, _TREE[1], ges + '.npz'), tsdf=volumes.astype(f64), max_l=sides.astype(f64),
             mid_p=centres.astype(f64))
    if JOINT_NOR:
        joints = joints.reshape(count, 21, 3)
        nor = (joints.astype(f64) - centres.astype(f64)[:, None, :]) / sides.astype(f64)[:, None, None] + 0.5
        np.save(os.path.join(out_dir, 'joint_nor', ges + '.npy'), nor.reshape(count, 63))
    np.save(os.path.join(out_dir, _TREE[2], ges + '.npy'), joints)
    np.save(os.path.join(out_dir, _TREE[3], ges + '.npy'), count)


def _gesture_with_aug(engine, records, joints):
    """One gesture of the fork's frame loop with AUG (cut_version/pre/read_MSRA.py:82-102).  Per frame the fork
    consumes the global np.random stream in this order: the resampling indices of get_points
    (cut_version/pre/process.py:52-67), then data_aug's two uniform and three randint draws (:76-94) - so the
    index vectors are drawn frame by frame here (by the library's NumPy-compatible generator) with the
    augmentation parameters in between; everything that does not touch the stream is batched: gather +
    extrema + voxel loop of all frames on the GPU, the fork's float64 linear algebra per frame on the host,
    and one more batched voxeliser call for the augmented clouds' bounding boxes."""
    import process
    n = len(records)
    n_valid = engine.process_begin(records)
    idx = np.empty((n, POINTS_NUM), dtype=np.int32)
    params = []
    for i in range(n):
        idx[i] = _pkg.capi.resample_indices(n_valid[i:i + 1], POINTS_NUM)[0]
        params.append(process.draw_aug())
    clouds, volumes, sides, centres = engine.process_finish(idx, POINTS_NUM)
    gt3 = joints.reshape(n, 21, 3)
    clouds_aug = np.empty([n, POINTS_NUM, 3])
    gt_aug = np.empty(gt3.shape)
    for i in range(n):
        clouds_aug[i], gt_aug[i] = process.apply_aug(params[i], gt3[i], clouds[i])
    # tsdf_f(POINT_CLOUD_AUG[i]): the bounding box of the augmented cloud (max_min_point,
    # cut_version/pre/process.py:153-170: z ignores exact zeros, rounded to float32) over the ORIGINAL depth map
    extrema = np.empty((n, 6), dtype=np.float32)
    for i in range(n):
        x, y, z = clouds_aug[i, :, 0], clouds_aug[i, :, 1], clouds_aug[i, :, 2]
        z = z[z != 0]
        extrema[i] = (np.max(x), np.max(y), np.max(z), np.min(x), np.min(y), np.min(z))
    vol_aug, sides_aug, centres_aug = engine.voxelize_frames(records, layout=_pkg.LAYOUT_CXYZ,
                                                             bbox_mode=_pkg.BBOX_EXTREMA, aux=extrema)
    return (clouds, volumes, sides, centres), (clouds_aug, vol_aug, sides_aug, centres_aug, gt_aug)


def _store_gesture_aug(out_dir, ges, count, clouds_aug, vol_aug, sides_aug, centres_aug, gt_aug):
    """The _aug files of a gesture (cut_version/pre/read_MSRA.py:115-124)."""
    f64 = np.float64
    np.save(os.path.join(out_dir, 'Point_Cloud_aug', ges + '.npy'), clouds_aug)
    np.savez(os.path.join(out_dir, 'TSDF_aug', ges + '.npz'), tsdf=vol_aug.astype(f64), max_l=sides_aug.astype(f64),
             mid_p=centres_aug.astype(f64))
    np.save(os.path.join(out_dir, 'ground_truth_aug', ges + '.npy'), gt_aug)
    np.save(os.path.join(out_dir, 'num_aug', ges + '.npy'), count)
    # the fork REBINDS its whole joint_nor_aug array to the row of the frame at hand (:101-102, no [i]): what it
    # saves is the [1,63] row of the gesture's last frame
    last = (gt_aug[-1] - centres_aug.astype(f64)[-1]) / sides_aug.astype(f64)[-1] + 0.5
    np.save(os.path.join(out_dir, 'joint_nor_aug', ges + '.npy'), last.reshape(-1, 63))


def main():
    "main part"
    if AUG and not JOINT_NOR:
        raise NotImplementedError(
            "AUG = True with the original's tree: the original's data_aug raises AxisError (pre/process.py:252); "
            "set JOINT_NOR = True for the later fork's tree with augmentation (cut_version/pre/read_MSRA.py), or "
            "see the on-device augmentation (SURVEY.md section 8f, N2)")
    _ensure_dir(SAVE_dir, "result")
    engine = _pkg.default_voxelizer()
    people = sorted(os.listdir(DB_dir))                      # 9 subjects ...
    if JOINT_NOR:
        people = people[:3]                                  # ... of which the fork takes three (:27)
    poses = sorted(os.listdir(os.path.join(DB_dir, people[0])))   # ... x 17 gestures in the full data base
    tree = _TREE + (('joint_nor',) if JOINT_NOR else ())
    if AUG:
        tree = tree + tuple(name + '_aug' for name in tree)
    for who in people:
        out_dir = os.path.join(SAVE_dir, who)
        _ensure_dir(out_dir, who)
        for name in tree:
            _ensure_dir(os.path.join(out_dir, name))
        frames_of_subject = 0
        for ges in poses:
            count, joints, records = _load_gesture(os.path.join(DB_dir, who, ges))
            frames_of_subject += count
            if AUG:
                plain, aug = _gesture_with_aug(engine, records, joints)
                _store_gesture(out_dir, ges, count, joints, *plain)
                _store_gesture_aug(out_dir, ges, count, *aug)
            else:
                # one batched call per gesture instead of DataProcess(...).process() per file (:98-111)
                clouds, volumes, sides, centres = engine.process_frames(records, POINTS_NUM)
                _store_gesture(out_dir, ges, count, joints, clouds, volumes, sides, centres)
            print('%s-%s files saved.' % (who, ges))
        if not JOINT_NOR:
            np.save(os.path.join(SAVE_dir, 'data_num-%s.npy' % who), frames_of_subject)
        print('%s total number saved.' % who)


def read_joint(f_dir):
    """joint.txt of a gesture folder: first line = number of .bin files, then 63 ground-truth values
    per frame.  Returns (bin_num, float32[bin_num,63])."""
    path = os.path.join(f_dir, 'joint.txt')
    with open(path) as fh:
        bin_num = int(fh.readline().strip())
    return bin_num, np.loadtxt(path, dtype=np.float32, skiprows=1)


def read_bin(f_name):
    """One MSRA depth file: int32[6] = img_w, img_h, left, top, right, bottom, then the float32 crop to
    the end of the file.  Returns (header, depth)."""
    raw = np.fromfile(f_name, dtype=np.uint8)
    head = raw[:24 - (min(raw.size, 24) % 4) if raw.size < 24 else 24]
    body = raw[24:24 + 4 * ((max(raw.size, 24) - 24) // 4)]          # whole floats only, like np.fromfile
    return head[:4 * (head.size // 4)].view(np.int32).copy(), body.view(np.float32).copy()


if __name__ == '__main__':
    main()
