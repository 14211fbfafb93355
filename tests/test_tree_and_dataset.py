"""The callers either side of the hot path against the golden
``result/`` tree written by the UNMODIFIED reference (tests/golden/make_golden_tree.py ->
tests/golden/msra_tree.npz): the ``3D_CNN/dataset.py`` consumer (CPU) and the ``pre/read_MSRA.py``
preprocessing driver (GPU)."""
import hashlib
import importlib
import importlib.util
import os
import sys

import numpy as np
import pytest

from helpers import graft, pkg

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_tree as mini  # noqa: E402   (only its data-base writer is used here)

CNN_DIR = os.path.join(graft.PKG_DIR, "3D_CNN")
PRE_DIR = os.path.join(graft.PKG_DIR, "pre")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "msra_tree.npz"))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _import(name, directory):
    if directory not in sys.path:
        sys.path.insert(0, directory)
    return importlib.import_module(name)


def _write_tree(golden, root, gt_3d):
    """The reference's result/ tree (pre/read_MSRA.py:119-126), rebuilt from the fixture."""
    for sub in mini.SUBJECTS:
        for name in ("Point_Cloud", "TSDF", "ground_truth", "num"):
            os.makedirs(os.path.join(root, sub, name))
        for ges in mini.GESTURES:
            k = "%s_%s_" % (sub, ges)
            np.savez(os.path.join(root, sub, "TSDF", "%s.npz" % ges), tsdf=golden[k + "tsdf"].astype(np.float64),
                     max_l=golden[k + "max_l"], mid_p=golden[k + "mid_p"])
            gt = golden[k + "gt"]
            np.save(os.path.join(root, sub, "ground_truth", "%s.npy" % ges), gt.reshape(-1, 21, 3) if gt_3d else gt)
            np.save(os.path.join(root, sub, "num", "%s.npy" % ges), golden[k + "num"])
        np.save(os.path.join(root, "data_num-%s.npy" % sub), golden["data_num_%s" % sub])


# ----------------------------------------------------------------------------- consumer (CPU)
@pytest.mark.parametrize("gt_3d", [True, False])
def test_dataset_matches_reference_loader(golden, tmp_path, gt_3d):
    """MSRA_Dataset: 3D_CNN/dataset.py:16-117.  gt_3d=False is the tree pre/read_MSRA.py really
    writes ([n,63] ground truth), which the reference loader cannot read (NameError, :105-108)."""
    dataset = _import("dataset", CNN_DIR)
    root = str(tmp_path / "result")
    _write_tree(golden, root, gt_3d)
    for train, tag in ((True, "train"), (False, "test")):
        ds = dataset.MSRA_Dataset(root, None, train=train, aug=False)
        assert len(ds) == int(golden["ds_%s_len" % tag])
        assert ds.tsdf.dtype == np.float32 and ds.tsdf.shape[1:] == (3, 32, 32, 32)
        assert _sha(ds.tsdf) == str(golden["ds_%s_tsdf_sha256" % tag])          # same frames, same order
        assert np.array_equal(ds.ground_truth, golden["ds_%s_gt" % tag])        # z flipped (:107)
        assert np.array_equal(ds.max_l, golden["ds_%s_max_l" % tag])
        assert np.array_equal(ds.mid_p, golden["ds_%s_mid_p" % tag])
        item = ds[len(ds) - 1]
        assert len(item) == int(golden["ds_%s_item_len" % tag]) == 4
        assert item[0].shape == (3, 32, 32, 32) and item[1].shape == (63,) and item[3].shape == (3,)
    # consumed by a DataLoader like 3D_CNN/train.py:87-91
    import torch
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False, num_workers=0)
    tsdf, gt, max_l, mid_p = next(iter(loader))
    assert tuple(tsdf.shape) == (4, 3, 32, 32, 32) and tsdf.dtype == torch.float32
    assert tuple(gt.shape) == (4, 63) and tuple(max_l.shape) == (4,) and tuple(mid_p.shape) == (4, 3)


def test_dataset_cut_version_variant(golden, tmp_path):
    """variant='cut': the later fork's loader (cut_version/3D_CNN/dataset.py:16-108): 3 subjects x 3
    gestures, ground truth [21,3] unflipped, joint_nor from the tree as a fifth member.  Where the
    reference tree is mounted (the build container) the unmodified fork loads the same tree for
    comparison; elsewhere the expectation is rebuilt from the fixture."""
    import types
    dataset = _import("dataset", CNN_DIR)
    root = str(tmp_path / "result")
    _write_tree(golden, root, True)
    want_jn = {}
    for sub in mini.SUBJECTS:
        os.makedirs(os.path.join(root, sub, "joint_nor"))
        for ges in mini.GESTURES:
            k = "%s_%s_" % (sub, ges)
            gt = golden[k + "gt"].reshape(-1, 21, 3).astype(np.float64)
            jn = np.stack([(gt[i] - golden[k + "mid_p"][i]) / golden[k + "max_l"][i] + 0.5 for i in range(gt.shape[0])])
            np.save(os.path.join(root, sub, "joint_nor", "%s.npy" % ges), jn.reshape(-1, 63))
            want_jn[(sub, ges)] = jn.reshape(-1, 63).astype(np.float32)
    opt = types.SimpleNamespace(size="small", test_index=1, PCA_SZ=63)
    ref_mod = None
    ref_dir = "/root/reference/cut_version/3D_CNN"
    if os.path.isdir(ref_dir):
        spec = importlib.util.spec_from_file_location("ref_cut_dataset", os.path.join(ref_dir, "dataset.py"))
        ref_mod = importlib.util.module_from_spec(spec)
        keep = {k: os.environ.get(k) for k in ("CUDA_DEVICE_ORDER", "CUDA_VISIBLE_DEVICES")}
        spec.loader.exec_module(ref_mod)                       # (it overwrites CUDA_VISIBLE_DEVICES at import)
        for k, v in keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for train in (True, False):
        ds = dataset.MSRA_Dataset(root, opt, train=train, aug=False, variant="cut")
        subs = [s for i, s in enumerate(mini.SUBJECTS[:3]) if (i != 1) == train]
        keys = [(s, g) for s in subs for g in mini.GESTURES[:3]]
        assert len(ds) == sum(want_jn[k].shape[0] for k in keys)
        assert np.array_equal(ds.joint_nor, np.concatenate([want_jn[k] for k in keys]))
        want_gt = np.concatenate([golden["%s_%s_gt" % k].reshape(-1, 21, 3) for k in keys])
        assert np.array_equal(ds.ground_truth, want_gt)                         # [n,21,3], no z flip
        item = ds[0]
        assert len(item) == 5 and item[1].shape == (21, 3) and item[4].shape == (63,)
        if ref_mod is not None:
            rd = ref_mod.MSRA_Dataset(root, opt, train=train, aug=False)
            assert len(rd) == len(ds)
            for name in ("tsdf", "ground_truth", "max_l", "mid_p", "joint_nor"):
                assert np.array_equal(getattr(rd, name), getattr(ds, name)), name
            for a, b in zip(rd[len(rd) - 1], ds[len(ds) - 1]):
                assert np.array_equal(a, b)


# ----------------------------------------------------------------------------- preprocessing driver (GPU)
@pytest.mark.gpu
def test_read_msra_main_writes_the_reference_tree(golden, tmp_path):
    """pre/read_MSRA.py main(): same files, same arrays, bit for bit, as the unmodified reference on
    the same synthetic MSRA data base and np.random seed."""
    read_MSRA = _import("read_MSRA", PRE_DIR)
    db = str(tmp_path / "db") + os.sep
    os.makedirs(db)
    mini.write_mini_db(db)
    h, d = read_MSRA.read_bin(os.path.join(db, "P0", "1", "000000_depth.bin"))      # pre/read_MSRA.py:155-164
    fr = pkg.synth.make_frame(mini.frames_of(0, 0)[0])
    assert h.dtype == np.int32 and np.array_equal(h, fr["header"]) and np.array_equal(d, fr["depth"])
    n, gt = read_MSRA.read_joint(os.path.join(db, "P0", "1"))                         # :143-152
    assert n == len(mini.frames_of(0, 0)) and gt.dtype == np.float32 and gt.shape == (n, 63)

    res = str(tmp_path / "result")
    read_MSRA.DB_dir, read_MSRA.SAVE_dir, read_MSRA.AUG = db, res, False
    np.random.seed(int(golden["seed"]))
    read_MSRA.main()
    for sub in mini.SUBJECTS:
        assert np.array_equal(np.load(os.path.join(res, "data_num-%s.npy" % sub)), golden["data_num_%s" % sub])
        for ges in mini.GESTURES:
            k = "%s_%s_" % (sub, ges)
            z = np.load(os.path.join(res, sub, "TSDF", "%s.npz" % ges))
            assert z["tsdf"].dtype == np.float64 and z["max_l"].dtype == np.float64
            assert np.array_equal(z["tsdf"], golden[k + "tsdf"].astype(np.float64)), k
            assert np.array_equal(np.signbit(z["tsdf"]), np.signbit(golden[k + "tsdf"]))
            assert np.array_equal(z["max_l"], golden[k + "max_l"]) and np.array_equal(z["mid_p"], golden[k + "mid_p"])
            pc = np.load(os.path.join(res, sub, "Point_Cloud", "%s.npy" % ges))
            assert pc.dtype == np.float64 and _sha(pc) == str(golden[k + "pc_sha256"]), k
            assert np.array_equal(pc[:, :32, :], golden[k + "pc_head"])
            assert np.array_equal(np.load(os.path.join(res, sub, "ground_truth", "%s.npy" % ges)), golden[k + "gt"])
            assert np.array_equal(np.load(os.path.join(res, sub, "num", "%s.npy" % ges)), golden[k + "num"])
    read_MSRA.AUG = True
    with pytest.raises(NotImplementedError):        # the original's tree with AUG: the original raises too
        read_MSRA.main()
    read_MSRA.AUG = False

    # the later fork's tree (cut_version/pre/read_MSRA.py:60-113): first three subjects, same volumes, ground
    # truth [n,21,3], joint_nor/<ges>.npy in float64, no data_num file; MSRA_Dataset(variant='cut') loads it
    res2 = str(tmp_path / "result_cut")
    read_MSRA.SAVE_dir, read_MSRA.JOINT_NOR = res2, True
    np.random.seed(int(golden["seed"]))
    try:
        read_MSRA.main()
    finally:
        read_MSRA.JOINT_NOR = False
    assert not os.path.exists(os.path.join(res2, "data_num-P0.npy")) and not os.path.exists(os.path.join(res2, "P3"))
    for sub in mini.SUBJECTS[:3]:
        for ges in mini.GESTURES:
            k = "%s_%s_" % (sub, ges)
            z = np.load(os.path.join(res2, sub, "TSDF", "%s.npz" % ges))
            assert np.array_equal(z["tsdf"], golden[k + "tsdf"].astype(np.float64))
            gt = np.load(os.path.join(res2, sub, "ground_truth", "%s.npy" % ges))
            assert gt.shape[1:] == (21, 3) and np.array_equal(gt.reshape(-1, 63), golden[k + "gt"])
            jn = np.load(os.path.join(res2, sub, "joint_nor", "%s.npy" % ges))
            want = np.stack([(gt[i].astype(np.float64) - z["mid_p"][i]) / z["max_l"][i] + 0.5 for i in range(gt.shape[0])])
            assert jn.dtype == np.float64 and np.array_equal(jn, want.reshape(-1, 63))
    import types
    dataset = _import("dataset", CNN_DIR)
    ds = dataset.MSRA_Dataset(res2, types.SimpleNamespace(size="small", test_index=1, PCA_SZ=63), train=True, variant="cut")
    assert len(ds[0]) == 5 and ds.joint_nor.shape == (len(ds), 63)


@pytest.mark.gpu
def test_dataset_from_database_on_device(golden, tmp_path):
    """MSRA_Dataset.from_database: data base -> GPU-resident dataset with no result/ tree in between;
    with the reference's seed it holds exactly what the reference loader returns."""
    import torch
    dataset = _import("dataset", CNN_DIR)
    db = str(tmp_path / "db") + os.sep
    os.makedirs(db)
    mini.write_mini_db(db)
    # the reference preprocesses EVERY subject in order (pre/read_MSRA.py:51), so the np.random stream
    # a subject sees depends on the ones before it: replay the full order, keep the test split
    np.random.seed(int(golden["seed"]))
    full = {}
    for s, sub in enumerate(mini.SUBJECTS):
        ds = dataset.MSRA_Dataset.from_database(db, train=False, device="cuda", test_idx=s)
        full[s] = ds
    ds = full[2]                                                          # the reference's test_idx (:21)
    assert ds.tsdf.is_cuda and len(ds) == int(golden["ds_test_len"])
    assert _sha(ds.tsdf.cpu().numpy()) == str(golden["ds_test_tsdf_sha256"])
    assert np.array_equal(ds.ground_truth.cpu().numpy(), golden["ds_test_gt"])
    assert np.array_equal(ds.max_l.cpu().numpy(), golden["ds_test_max_l"])
    assert np.array_equal(ds.mid_p.cpu().numpy(), golden["ds_test_mid_p"])
    tsdf, gt, max_l, mid_p = ds[torch.arange(0, 4, device="cuda")]        # a batch without a host round trip
    assert tsdf.is_cuda and tuple(tsdf.shape) == (4, 3, 32, 32, 32)
    # deterministic variant: grid from all valid pixels (pre/tsdf_numba.py semantics), one fused launch per gesture
    dp = dataset.MSRA_Dataset.from_database(db, train=True, device="cuda", bbox="pixels")
    assert len(dp) == int(golden["ds_train_len"]) and float(dp.tsdf.abs().max()) <= 1.0
    # the same set kept PACKED in host memory (dtsdf_voxelize_host_packed), samples expanded on demand
    lazy = dataset.MSRA_Dataset.from_database(db, train=True, device=None, bbox="pixels", packed_host=True)
    assert len(lazy) == len(dp) and tuple(lazy.tsdf.shape) == tuple(dp.tsdf.shape)
    assert lazy.tsdf.nbytes < dp.tsdf.numel() * 4 / 4
    dense = dp.tsdf.cpu().numpy()
    for i in (0, 1, len(lazy) // 2, len(lazy) - 1):
        t, g, ml, mp = lazy[i]
        assert t.dtype == np.float32 and np.array_equal(t.view(np.uint32), dense[i].view(np.uint32))
        assert ml == dp.max_l[i].item() and np.array_equal(g, dp.ground_truth[i].cpu().numpy())
    loader = torch.utils.data.DataLoader(lazy, batch_size=5, shuffle=False)
    got = torch.cat([b[0] for b in loader])
    assert torch.equal(got, dp.tsdf.cpu())


@pytest.mark.gpu
def test_read_msra_main_writes_the_forks_tree_with_augmentation(tmp_path):
    """pre/read_MSRA.py with JOINT_NOR and AUG: every file of the later fork's result/ tree - plain half and
    _aug half - bit-identical to the UNMODIFIED fork's (cut_version/pre/read_MSRA.py:17-140 with AUG = True,
    cut_version/pre/process.py:52-134; tests/golden/make_golden_tree_cut.py), np.random stream position
    afterwards included.  DataProcess.data_aug(joints, point_clouds) itself against the frozen clouds."""
    gold = np.load(os.path.join(HERE, "golden", "msra_tree_cut.npz"))
    read_MSRA = _import("read_MSRA", PRE_DIR)
    process = _import("process", PRE_DIR)
    db = str(tmp_path / "db") + os.sep
    os.makedirs(db)
    mini.write_mini_db(db)
    res = str(tmp_path / "result")
    read_MSRA.DB_dir, read_MSRA.SAVE_dir, read_MSRA.AUG, read_MSRA.JOINT_NOR = db, res, True, True
    np.random.seed(int(gold["seed"]))
    try:
        read_MSRA.main()
    finally:
        read_MSRA.AUG, read_MSRA.JOINT_NOR = False, False
    assert np.array_equal(np.random.randint(0, 1 << 30, size=4), gold["stream_tail"])
    assert sorted(os.listdir(res)) == [str(s) for s in gold["subjects"]]
    for sub in gold["subjects"]:
        assert sorted(os.listdir(os.path.join(res, str(sub)))) == sorted(
            n + sfx for n in ("Point_Cloud", "TSDF", "ground_truth", "num", "joint_nor") for sfx in ("", "_aug"))
        for ges in gold["gestures"]:
            for sfx in ("", "_aug"):
                k = "%s_%s%s_" % (sub, ges, sfx)
                z = np.load(os.path.join(res, str(sub), "TSDF" + sfx, "%s.npz" % ges))
                assert z["tsdf"].dtype == np.float64 and np.array_equal(z["tsdf"], gold[k + "tsdf"].astype(np.float64)), k
                assert np.array_equal(np.signbit(z["tsdf"]), np.signbit(gold[k + "tsdf"]))
                assert z["max_l"].dtype == np.float64 and np.array_equal(z["max_l"], gold[k + "max_l"]), k
                assert np.array_equal(z["mid_p"], gold[k + "mid_p"]), k
                pc = np.load(os.path.join(res, str(sub), "Point_Cloud" + sfx, "%s.npy" % ges))
                assert pc.dtype == np.float64 and np.array_equal(pc[:, :32, :], gold[k + "pc_head"]), k
                assert _sha(pc) == str(gold[k + "pc_sha256"]), k
                for name in ("ground_truth", "num", "joint_nor"):
                    got = np.load(os.path.join(res, str(sub), name + sfx, "%s.npy" % ges))
                    want = gold[k + {"ground_truth": "gt", "num": "num", "joint_nor": "joint_nor"}[name]]
                    assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (k, name)
    # the drop-in method itself: DataProcess(data, point_num, aug).data_aug(joints, point_clouds)
    fr = pkg.synth.make_frame(5)
    dp = process.DataProcess(fr, 6000, True)
    np.random.seed(3)
    pc = dp.get_points()
    pc_aug, gt_aug = dp.data_aug(np.arange(63, dtype=np.float32), pc)
    assert pc_aug.shape == (6000, 3) and pc_aug.dtype == np.float64 and gt_aug.shape == (21, 3)
    tsdf_aug, max_l_aug, mid_p_aug = dp.tsdf_f(pc_aug)
    assert tsdf_aug.shape == (3, 32, 32, 32) and max_l_aug > 0
    with pytest.raises(NotImplementedError):
        dp.data_aug(pc)                                     # the original's signature: the original raises AxisError
