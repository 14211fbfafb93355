#!/usr/bin/env python
"""Freeze the ``result/`` tree of the later fork, augmentation included, by EXECUTING THE UNMODIFIED FORK in
this container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden_tree_cut.py        ->  tests/golden/msra_tree_cut.npz

cut_version/pre/read_MSRA.py main() (:17-140) with ``AUG = True`` on the synthetic mini MSRA data base of
make_golden_tree.py (the fork takes the first 3 subjects, :27), ``np.random.seed(SEED)`` first.  Per frame
the fork draws, in this order, from the global np.random stream: the resampling indices of get_points
(cut_version/pre/process.py:52-67), then data_aug's two ``uniform`` and three ``randint`` (:69-98).  What is
frozen: every array of both halves of the tree - TSDF / Point_Cloud / ground_truth / num / joint_nor and their
``_aug`` twins (``joint_nor_aug`` is the [1,63] row of the LAST frame: the fork rebinds the whole array at
cut_version/pre/read_MSRA.py:101-102).
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_tree as mini  # noqa: E402

SEED = 20240608
SUBJECTS = mini.SUBJECTS[:3]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    work = tempfile.mkdtemp(prefix="msra_golden_cut_")
    db = os.path.join(work, "db") + os.sep
    res = os.path.join(work, "result")
    os.makedirs(db)
    mini.write_mini_db(db)
    sys.path.insert(0, "/root/reference/cut_version/pre")
    import read_MSRA
    assert read_MSRA.__file__.startswith("/root/reference/cut_version/")
    read_MSRA.DB_dir = db
    read_MSRA.SAVE_dir = res
    read_MSRA.AUG = True
    np.random.seed(SEED)
    read_MSRA.main()
    tail = np.random.randint(0, 1 << 30, size=4)            # where the stream stands afterwards

    out = {"seed": np.int64(SEED), "subjects": np.array(SUBJECTS), "gestures": np.array(mini.GESTURES),
           "stream_tail": tail}
    for sub in SUBJECTS:
        assert not os.path.exists(os.path.join(res, "data_num-%s.npy" % sub))
        for ges in mini.GESTURES:
            for sfx in ("", "_aug"):
                key = "%s_%s%s_" % (sub, ges, sfx)
                z = np.load(os.path.join(res, sub, "TSDF" + sfx, "%s.npz" % ges))
                tsdf = z["tsdf"]
                assert tsdf.dtype == np.float64 and np.array_equal(tsdf, tsdf.astype(np.float32).astype(np.float64))
                out[key + "tsdf"] = tsdf.astype(np.float32)
                out[key + "max_l"] = z["max_l"]
                out[key + "mid_p"] = z["mid_p"]
                pc = np.load(os.path.join(res, sub, "Point_Cloud" + sfx, "%s.npy" % ges))
                assert pc.dtype == np.float64
                out[key + "pc_sha256"] = np.array(sha(pc))
                out[key + "pc_head"] = pc[:, :32, :].copy()
                out[key + "gt"] = np.load(os.path.join(res, sub, "ground_truth" + sfx, "%s.npy" % ges))
                out[key + "num"] = np.load(os.path.join(res, sub, "num" + sfx, "%s.npy" % ges))
                out[key + "joint_nor"] = np.load(os.path.join(res, sub, "joint_nor" + sfx, "%s.npy" % ges))
    dst = os.path.join(HERE, "msra_tree_cut.npz")
    np.savez_compressed(dst, **out)
    shutil.rmtree(work)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
