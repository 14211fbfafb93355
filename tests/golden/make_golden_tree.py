#!/usr/bin/env python
"""Freeze golden vectors of the callers either side of the hot path by EXECUTING THE UNMODIFIED
REFERENCE in this container (needs /root/reference; never runs on the GPU box):

    python tests/golden/make_golden_tree.py        ->  tests/golden/msra_tree.npz

  * pre/read_MSRA.py main() (:37-140) on a synthetic mini MSRA data base (4 subjects x 5 gestures,
    2-3 frames each, written by ``write_mini_db`` below in the real on-disk format:
    ``%06d_depth.bin`` = 6 x int32 + float32 crop, ``joint.txt`` = count line + 63 floats per frame),
    with ``np.random.seed(SEED)`` set first: the whole ``result/`` tree it writes;
  * 3D_CNN/dataset.py MSRA_Dataset (:16-117) on that tree, train and test split.  As shipped it
    cannot load the tree its own preprocessing writes (``g_t`` is only assigned for 3-D ground truth,
    :105-108, and read_MSRA saves 2-D), so the ground-truth files are re-saved as [n,21,3] first -
    the shape the loader was written for; what is frozen is the frame ORDER and the ground truth it
    returns (z flipped, :107).
"""
import hashlib
import io
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "handposeestimation-with-3d-cnns_b200"))
import synth  # noqa: E402

SEED = 20240607
SUBJECTS = ["P0", "P1", "P2", "P3"]
GESTURES = ["1", "2", "3", "4", "5"]


def frames_of(si, gi):
    """Seeds of the synthetic frames of subject si, gesture gi (2 or 3 frames: np.loadtxt of a single row is 1-D and the reference indexes [i, :])."""
    n = 2 + (si + gi) % 2
    return [1000 + 100 * si + 10 * gi + k for k in range(n)]


def joints_of(si, gi, n):
    rng = np.random.default_rng(7000 + 10 * si + gi)
    return rng.uniform(-120.0, 120.0, size=(n, 63)).astype(np.float32)


def write_mini_db(db_dir):
    """The MSRA on-disk layout read by pre/read_MSRA.py:47-48,80,99-102."""
    for si, sub in enumerate(SUBJECTS):
        for gi, ges in enumerate(GESTURES):
            d = os.path.join(db_dir, sub, ges)
            os.makedirs(d)
            seeds = frames_of(si, gi)
            gt = joints_of(si, gi, len(seeds))
            with open(os.path.join(d, "joint.txt"), "w") as f:
                f.write("%d\n" % len(seeds))
                for row in gt:
                    f.write(" ".join("%.6f" % v for v in row) + "\n")
            for k, seed in enumerate(seeds):
                fr = synth.make_frame(seed)
                with open(os.path.join(d, "%06d_depth.bin" % k), "wb") as f:
                    f.write(fr["header"].astype(np.int32).tobytes())
                    f.write(fr["depth"].astype(np.float32).tobytes())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    work = tempfile.mkdtemp(prefix="msra_golden_")
    db = os.path.join(work, "db") + os.sep
    res = os.path.join(work, "result")
    os.makedirs(db)
    write_mini_db(db)

    # ---- the reference's offline preprocessing, unmodified
    sys.path.insert(0, "/root/reference/pre")
    import read_MSRA
    read_MSRA.DB_dir = db
    read_MSRA.SAVE_dir = res
    read_MSRA.AUG = False
    np.random.seed(SEED)
    read_MSRA.main()

    out = {"seed": np.int64(SEED), "subjects": np.array(SUBJECTS), "gestures": np.array(GESTURES)}
    for si, sub in enumerate(SUBJECTS):
        out["data_num_%s" % sub] = np.load(os.path.join(res, "data_num-%s.npy" % sub))
        for gi, ges in enumerate(GESTURES):
            key = "%s_%s_" % (sub, ges)
            z = np.load(os.path.join(res, sub, "TSDF", "%s.npz" % ges))
            tsdf = z["tsdf"]
            assert tsdf.dtype == np.float64 and np.array_equal(tsdf, tsdf.astype(np.float32).astype(np.float64))
            out[key + "tsdf"] = tsdf.astype(np.float32)           # every stored value is a float32 value
            out[key + "max_l"] = z["max_l"]
            out[key + "mid_p"] = z["mid_p"]
            pc = np.load(os.path.join(res, sub, "Point_Cloud", "%s.npy" % ges))
            assert pc.dtype == np.float64
            out[key + "pc_sha256"] = np.array(sha(pc))
            out[key + "pc_head"] = pc[:, :32, :].copy()
            out[key + "gt"] = np.load(os.path.join(res, sub, "ground_truth", "%s.npy" % ges))
            out[key + "num"] = np.load(os.path.join(res, sub, "num", "%s.npy" % ges))

    # ---- the reference's consumer, unmodified, on the tree with [n,21,3] ground truth
    for sub in SUBJECTS:
        for ges in GESTURES:
            p = os.path.join(res, sub, "ground_truth", "%s.npy" % ges)
            np.save(p, np.load(p).reshape(-1, 21, 3))
    sys.path.insert(0, "/root/reference/3D_CNN")
    keep = {k: os.environ.get(k) for k in ("CUDA_DEVICE_ORDER", "CUDA_VISIBLE_DEVICES")}
    import dataset as ref_dataset          # sets CUDA_VISIBLE_DEVICES at import (3D_CNN/dataset.py:12-13)
    for k, v in keep.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    for train in (True, False):
        ds = ref_dataset.MSRA_Dataset(res, None, train=train, aug=False)
        tag = "train" if train else "test"
        out["ds_%s_len" % tag] = np.int64(len(ds))
        out["ds_%s_gt" % tag] = ds.ground_truth.copy()
        out["ds_%s_max_l" % tag] = ds.max_l.copy()
        out["ds_%s_mid_p" % tag] = ds.mid_p.copy()
        out["ds_%s_tsdf_sha256" % tag] = np.array(sha(ds.tsdf))
        item = ds[len(ds) - 1]
        out["ds_%s_item_len" % tag] = np.int64(len(item))
    np.savez_compressed(os.path.join(HERE, "msra_tree.npz"), **out)
    shutil.rmtree(work)
    print("wrote", os.path.join(HERE, "msra_tree.npz"), os.path.getsize(os.path.join(HERE, "msra_tree.npz")), "bytes")


if __name__ == "__main__":
    main()
