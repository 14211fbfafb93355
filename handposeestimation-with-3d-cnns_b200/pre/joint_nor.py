"""Drop-in for the reference's ``pre/joint_nor.py``: joint normalisation into the voxel cube,
``(gt - mid_p) / max_l + 0.5``, evaluated by the CUDA library (``normalize_joints_kernel``).

  normalize(ground_truth, max_l, mid_p) ............................. pre/joint_nor.py:8-18
    ground_truth [N,21,3] (or [N,63]: the reference reshapes it but then fails to assign a
    [21,3] row into a [63] slot, :12-16; here both shapes work), max_l [N], mid_p [N,3]
    -> float64 array of the input's shape (``np.empty(gt_shap)``, :14).  Arithmetic is
    float32 on the GPU: within 1e-6 of the reference's float64 expression
    (tests/test_gpu_parity.py::test_joint_nor_module).
  read_joint(aug='') ................................................ pre/joint_nor.py:21-44
    walks ``DATA_dir`` (``./result``) and writes ``<subject>/joint_nor<aug>/<gesture>.npy``,
    the tree ``cut_version/3D_CNN/dataset.py:66,100-103`` loads.  The reference's own loop
    cannot run: ``'ground_truth%' % aug`` (:37) is a format error and ``np.save`` (:44) is
    handed the directory name; what it was written to do is specified here, not copied.
"""
import os
import os.path

import numpy as np

from _bind import package

_pkg = package()

DATA_dir = './result'


def normalize(ground_truth, max_l, mid_p):
    ground_truth = np.asarray(ground_truth)
    gt_shap = ground_truth.shape
    n = gt_shap[0] if len(gt_shap) > 1 else 1
    joints = np.ascontiguousarray(ground_truth, dtype=np.float32).reshape(n, 63)
    sides = np.ascontiguousarray(max_l, dtype=np.float32).reshape(n)
    centres = np.ascontiguousarray(mid_p, dtype=np.float32).reshape(n, 3)
    out = _pkg.default_voxelizer().normalize_joints(joints, sides, centres, clamp01=False)
    return out.astype(np.float64).reshape(gt_shap)


def read_joint(aug=''):
    subjects = sorted(os.listdir(DATA_dir))
    for sub in subjects:
        sub_dir = os.path.join(DATA_dir, sub)
        if not os.path.isdir(os.path.join(sub_dir, 'ground_truth%s' % aug)):
            continue
        out_dir = os.path.join(sub_dir, 'joint_nor%s' % aug)
        try:
            os.mkdir(out_dir)
        except OSError:
            print('failed create joint nor file')
        for name in sorted(os.listdir(os.path.join(sub_dir, 'ground_truth%s' % aug))):
            ges = os.path.splitext(name)[0]
            joint = np.load(os.path.join(sub_dir, 'ground_truth%s' % aug, name))
            datas = np.load(os.path.join(sub_dir, 'TSDF%s' % aug, ges + '.npz'))
            joint_nor = normalize(joint, datas['max_l'], datas['mid_p'])
            np.save(os.path.join(out_dir, '%s.npy' % ges), joint_nor)


if __name__ == "__main__":
    read_joint()
