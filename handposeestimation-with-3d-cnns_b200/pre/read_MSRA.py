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

Not reproduced: ``AUG = True`` (the reference's ``data_aug`` raises, pre/process.py:252) and the
trailing ``joint_pca(aug=False)`` of the reference's ``__main__`` block (pre/joint_pca.py is
self-declared untested and passes a function where data is expected, :34,39).
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

POINTS_NUM = 6000


def _mkdir(path, what):
    try:
        os.mkdir(path)
        print('create directory "%s".' % what)
    except OSError:
        print('directory "%s" already exist!' % what)


def main():
    "main part"
    if AUG:
        raise NotImplementedError(
            "AUG = True: the reference's data_aug raises AxisError (pre/process.py:252); "
            "augmentation is specified separately (SURVEY.md section 8f, N2)")
    _mkdir(SAVE_dir, "result")
    vox = _pkg.default_voxelizer()
    # there are 9 subjects and 17 gestures in each subject
    subject = sorted(os.listdir(DB_dir))
    gesture = sorted(os.listdir(os.path.join(DB_dir, subject[0])))
    for sub in subject:
        sub_dir = os.path.join(SAVE_dir, sub)
        _mkdir(sub_dir, sub)
        for name in ('Point_Cloud', 'TSDF', 'ground_truth', 'num'):
            try:
                os.mkdir(os.path.join(sub_dir, name))
            except OSError:
                pass

        total_num = 0
        for ges in gesture:
            file_dir = os.path.join(DB_dir, sub, ges)
            bin_num, ground_truth = read_joint(file_dir)
            total_num += bin_num

            frames = []
            for i in range(bin_num):
                header, depth = read_bin(os.path.join(file_dir, '%06d_depth.bin' % i))
                frames.append({'header': header, 'depth': depth})
            pcs, tsdf, max_l, mid_p = vox.process_frames(frames, POINTS_NUM)

            # the reference fills float64 arrays (np.empty default dtype, pre/read_MSRA.py:93-96)
            np.save(os.path.join(sub_dir, 'Point_Cloud', '%s.npy' % ges), pcs)
            np.savez(os.path.join(sub_dir, 'TSDF', '%s.npz' % ges),
                     tsdf=tsdf.astype(np.float64), max_l=max_l.astype(np.float64),
                     mid_p=mid_p.astype(np.float64))
            np.save(os.path.join(sub_dir, 'ground_truth', '%s.npy' % ges), ground_truth)
            np.save(os.path.join(sub_dir, 'num', '%s.npy' % ges), bin_num)
            print('%s-%s files saved.' % (sub, ges))
        np.save(os.path.join(SAVE_dir, 'data_num-%s.npy' % sub), total_num)
        print('%s total number saved.' % sub)


def read_joint(f_dir):
    "read the joint files and return the ground truth"
    # first line: number of bin files; the rest: 63 ground-truth values per frame
    f_name = os.path.join(f_dir, 'joint.txt')
    with open(f_name, 'r') as f:
        bin_num = int(f.readline())
    ground_truth = np.loadtxt(f_name, dtype=np.float32, skiprows=1)
    return bin_num, ground_truth


def read_bin(f_name):
    "read one bin file: 6 x int32 header, then the float32 depth crop to the end of the file"
    with open(f_name, 'rb') as f:
        header = np.fromfile(f, dtype=np.int32, count=6)
        depth = np.fromfile(f, dtype=np.float32)
    return header, depth


if __name__ == '__main__':
    main()
