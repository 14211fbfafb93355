"""Drop-in for the reference's ``3D_CNN/dataset.py``: ``MSRA_Dataset`` with the same constructor,
attributes, subject/gesture selection, frame order and ``__getitem__`` tuple - the consumer of the
D-TSDF volumes.

  MSRA_Dataset(root_path, opt, train=True, aug=False) ............... 3D_CNN/dataset.py:16-71
  __getitem__ -> (tsdf f32[3,32,32,32], ground_truth f32[63], max_l f32, mid_p f32[3]) ... :73-79
  __len__ ........................................................... :81-83

Two ways to fill it:
  * from the ``result/`` tree (the reference's way, :85-117): ``MSRA_Dataset(root_path, opt, ...)``;
  * straight from the MSRA data base, voxelised on the GPU with nothing written to disk:
    ``MSRA_Dataset.from_database(db_dir, ...)`` (the volumes never exist as float64 files).
``device='cuda'`` keeps the arrays in HBM as torch tensors, so ``__getitem__`` / a DataLoader batch
needs no per-step ``.cuda()`` copy (3D_CNN/train.py:232).

Deliberate differences from the reference:
  * ground truth saved as [n,63] (what pre/read_MSRA.py:123 writes) is accepted; the reference only
    assigns ``g_t`` for 3-D files (:105-108) and raises NameError on its own preprocessing output.
    Both shapes get the z flip of :107;
  * CUDA_VISIBLE_DEVICES is not overwritten at import (:12-13);
  * the PCA side data of ``size == 'full'`` (:165-181) is loaded only if ``./PCA`` exists.
``variant='cut'`` selects the behaviour of the later fork ``cut_version/3D_CNN/dataset.py`` (3 subjects,
ground truth [21,3], a fifth tuple member ``joint_nor``); see ``__init__``.
"""
import os
import os.path

import numpy as np
import torch
import torch.utils.data as data

from _bind import package


class MSRA_Dataset(data.Dataset):
    def __init__(self, root_path, opt=None, train=True, aug=False, device=None, size='small', test_idx=2,
                 variant='full'):
        self.root_path = root_path
        self.train = train
        self.size = size            # the reference hard-codes 'small' (:20)
        self.test_idx = test_idx    # and 2 (:21)
        self.PCA_SZ = 63
        self.AUG = aug
        self.device = device
        # variant='cut': the later, slimmed fork (cut_version/3D_CNN/dataset.py:16-66): size, test index
        # and PCA size come from ``opt`` (:20-22), 3 subjects, ground truth stays [n,21,3] without the
        # z flip (:95), and a fifth member joint_nor[63] (:66,100-103) - loaded from ``joint_nor<aug>/``
        # when the tree has it, otherwise computed on the GPU from ground truth, max_l and mid_p
        self.variant = variant
        if variant == 'cut' and opt is not None:
            self.size = getattr(opt, 'size', self.size)
            self.test_idx = getattr(opt, 'test_index', self.test_idx)
            self.PCA_SZ = getattr(opt, 'PCA_SZ', self.PCA_SZ)
        if variant == 'cut':
            self.SUBJECTS = 3
            self.GESTURES = 17 if self.size == 'full' else 3
        elif self.size == 'full':
            self.SUBJECTS = 9
            self.GESTURES = 17
        else:
            self.SUBJECTS = 4
            self.GESTURES = 5
        if root_path is not None:
            self._load_tree()
            self._finish()

    # ------------------------------------------------------------------ selection (:41-60, :119-163)
    def _subjects(self, names):
        """Indices of the subjects this split loads, in the reference's order."""
        if self.train:
            subs = [s for s in range(self.SUBJECTS) if s != self.test_idx]
        else:
            subs = [self.test_idx]
        return [names[s] for s in subs]

    def _load_tree(self):
        results = sorted(os.listdir(self.root_path))[:9]
        tsdf, gt, max_l, mid_p, jn = [], [], [], [], []
        cut = self.variant == 'cut'
        passes = [(self._subjects(results), '')]
        if self.AUG:
            passes.append(([results[s] for s in range(self.SUBJECTS)], '_aug'))     # :55-60
        for subs, suffix in passes:
            for sub in subs:
                data_dir = os.path.join(self.root_path, sub)
                print(("Training: " if self.train else "Testing: ") + data_dir)
                tsdf_file = sorted(os.listdir(os.path.join(data_dir, 'TSDF%s' % suffix)))
                g_file = sorted(os.listdir(os.path.join(data_dir, 'ground_truth%s' % suffix)))
                for ges in range(self.GESTURES):
                    z = np.load(os.path.join(data_dir, 'TSDF%s' % suffix, tsdf_file[ges]))
                    tsdf.append(z['tsdf'].astype(np.float32))
                    max_l.append(z['max_l'].astype(np.float32).reshape(-1))
                    mid_p.append(z['mid_p'].astype(np.float32).reshape(-1, 3))
                    g = np.load(os.path.join(data_dir, 'ground_truth%s' % suffix, g_file[ges])).astype(np.float32)
                    if not cut:
                        gt.append(self._flip_z(g))
                        continue
                    gt.append(g.reshape(-1, 21, 3))
                    n_dir = os.path.join(data_dir, 'joint_nor%s' % suffix)
                    if os.path.isdir(n_dir):
                        n_file = sorted(os.listdir(n_dir))
                        jn.append(np.load(os.path.join(n_dir, n_file[ges])).astype(np.float32).reshape(-1, 63))
                    else:
                        jn.append(package().default_voxelizer().normalize_joints(g.reshape(-1, 63), max_l[-1], mid_p[-1]))
        self.tsdf = np.concatenate(tsdf) if tsdf else np.empty((0, 3, 32, 32, 32), np.float32)
        self.ground_truth = np.concatenate(gt) if gt else np.empty((0, 63), np.float32)
        self.max_l = np.concatenate(max_l) if max_l else np.empty(0, np.float32)
        self.mid_p = np.concatenate(mid_p) if mid_p else np.empty((0, 3), np.float32)
        if cut:
            self.joint_nor = np.concatenate(jn) if jn else np.empty((0, 63), np.float32)
        self.total_num = self.tsdf.shape[0]

    @staticmethod
    def _flip_z(g):
        g = g.reshape(-1, 21, 3).copy()
        g[:, :, 2] = -g[:, :, 2]                      # :107
        return g.reshape(-1, 63)

    # ------------------------------------------------------------------ straight from the data base
    @classmethod
    def from_database(cls, db_dir, train=True, device='cuda', size='small', test_idx=2, point_num=6000,
                      bbox='points', packed_host=False):
        """Reads the ``.bin`` / ``joint.txt`` files of the selected subjects and gestures and
        voxelises them on the GPU; nothing is written to disk.

        ``bbox='points'`` is the reference's preprocessing (grid from the randomly resampled
        ``point_num``-point cloud, pre/process.py:13-28; consumes ``np.random``); ``bbox='pixels'``
        takes the grid from all valid pixels (pre/tsdf_numba.py:140-147; deterministic, and the
        whole frame is then one fused kernel).

        ``packed_host=True`` (with ``device=None`` and ``bbox='pixels'``): the volumes stay in host memory in
        the lossless packed form they crossed PCIe in (``PackedHostVolumeList``, ~47 KB instead of 393 KB per
        sample: the full MSRA set is 3.6 GB instead of the 30 GB of 3D_CNN/dataset.py:35) and
        ``__getitem__`` expands the sample it is asked for.
        """
        if packed_host and (device is not None or bbox != 'pixels'):
            raise ValueError("packed_host=True needs device=None and bbox='pixels'")
        pkg = package()
        import sys
        pre_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pre")
        if pre_dir not in sys.path:
            sys.path.insert(0, pre_dir)
        import read_MSRA
        self = cls(None, None, train=train, aug=False, device=device, size=size, test_idx=test_idx)
        subjects = sorted(os.listdir(db_dir))[:9]
        vox = pkg.default_voxelizer()
        tsdf, gt, max_l, mid_p = [], [], [], []
        for sub in self._subjects(subjects):
            gestures = sorted(os.listdir(os.path.join(db_dir, sub)))
            for ges in gestures[:self.GESTURES]:
                file_dir = os.path.join(db_dir, sub, ges)
                bin_num, ground_truth = read_MSRA.read_joint(file_dir)
                frames = []
                for i in range(bin_num):
                    header, depth = read_MSRA.read_bin(os.path.join(file_dir, '%06d_depth.bin' % i))
                    frames.append({'header': header, 'depth': depth})
                if bbox == 'points':
                    _, v, ml, mp = vox.process_frames(frames, point_num, want_points=False)
                elif packed_host:
                    v, ml, mp = vox.voxelize_frames_lazy(frames, layout=pkg.LAYOUT_CXYZ, bbox_mode=pkg.BBOX_PIXELS)
                else:
                    v, ml, mp = vox.voxelize_frames(frames, layout=pkg.LAYOUT_CXYZ, bbox_mode=pkg.BBOX_PIXELS)
                tsdf.append(v); max_l.append(ml); mid_p.append(mp)
                gt.append(self._flip_z(np.asarray(ground_truth, np.float32).reshape(bin_num, 63)))
        self.tsdf = pkg.PackedHostVolumeList(tsdf) if packed_host else np.concatenate(tsdf)
        self.ground_truth = np.concatenate(gt)
        self.max_l = np.concatenate(max_l)
        self.mid_p = np.concatenate(mid_p)
        self.total_num = self.tsdf.shape[0]
        self._finish()
        return self

    def _finish(self):
        if self.variant != 'cut' and self.size == 'full' and os.path.isdir('./PCA'):
            self.__loadPCA()
        if self.device is not None:
            dev = torch.device(self.device)
            self.tsdf = torch.from_numpy(self.tsdf).to(dev)
            self.ground_truth = torch.from_numpy(self.ground_truth).to(dev)
            self.max_l = torch.from_numpy(self.max_l).to(dev)
            self.mid_p = torch.from_numpy(self.mid_p).to(dev)
            if hasattr(self, 'ground_truth_pca'):
                self.ground_truth_pca = torch.from_numpy(self.ground_truth_pca).to(dev)
            if hasattr(self, 'joint_nor'):
                self.joint_nor = torch.from_numpy(self.joint_nor).to(dev)

    def __getitem__(self, index):
        """return index data"""
        if self.variant == 'cut':         # cut_version/3D_CNN/dataset.py:66
            return (self.tsdf[index], self.ground_truth[index], self.max_l[index], self.mid_p[index],
                    self.joint_nor[index])
        if hasattr(self, 'ground_truth_pca'):
            return (self.tsdf[index], self.ground_truth[index], self.max_l[index], self.mid_p[index],
                    self.ground_truth_pca[index])
        return self.tsdf[index], self.ground_truth[index], self.max_l[index], self.mid_p[index]

    def __len__(self):
        """to get all data number"""
        return self.tsdf.shape[0]

    def __loadPCA(self):
        all_f = os.listdir('./PCA')
        files = all_f[:9] if self.AUG else all_f[9:]
        z = np.load(os.path.join('./PCA', files[self.test_idx]))
        self.PCA_mean = z['pca_mean'].astype(np.float32)
        self.PCA_coeff = z['coeff'][:, 0:self.PCA_SZ].astype(np.float32)
        self.ground_truth_pca = np.dot(self.ground_truth - self.PCA_mean, self.PCA_coeff)
        self.PCA_coeff = torch.from_numpy(self.PCA_coeff)
        self.PCA_mean = torch.from_numpy(self.PCA_mean)
