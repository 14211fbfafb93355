"""The oracle (oracle/dtsdf_oracle.py) against the frozen outputs of the UNMODIFIED
reference (tests/golden/*.npz, made by tests/golden/make_golden.py).  Bit-exact."""
import numpy as np

from oracle import dtsdf_oracle as orc


def _frame(c, key="depth"):
    return {"header": c["header"], key: c["depth"]}


def test_point_cloud_matches_reference(forloop_golden):
    # pre/process.py:30-68
    for c in forloop_golden:
        pts = orc.point_cloud(c["header"], c["depth"])
        assert pts.dtype == np.float64
        assert pts.shape[0] == int(c["n_valid"])
        assert np.array_equal(pts[:64], c["pc_all_head"])
        assert np.array_equal(pts.sum(axis=0), c["pc_all_sum"])


def test_resample_consumes_global_rng_like_reference(forloop_golden):
    # pre/process.py:70-84 with the same np.random.seed as the golden run
    for c in forloop_golden:
        pts = orc.point_cloud(c["header"], c["depth"])
        np.random.seed(int(c["np_random_seed"]))
        pc = orc.set_length(pts, int(c["point_num"]))
        assert pc.shape == (int(c["point_num"]), 3)
        assert np.array_equal(pc[:64], c["pc_head"])
        if "pc" in c:
            assert np.array_equal(pc, c["pc"])
        pmax, pmin = orc.max_min_point(pc)
        assert pmax.dtype == np.float32
        assert np.array_equal(pmax, c["pmax"]) and np.array_equal(pmin, c["pmin"])


def test_grid_params_and_volume_match_for_loop(forloop_golden):
    # pre/process.py:86-100,122-200 == pre/tsdf_for.py:6-20,44-122
    for c in forloop_golden:
        vox_ori, voxel_len, trunc, max_l, mid = orc.grid_params(c["pmax"], c["pmin"])
        assert max_l.dtype == np.float32 and max_l == c["max_l"]
        assert np.array_equal(mid, c["mid_p"])
        vol = orc.tsdf_cal(c["header"], c["depth"], vox_ori, voxel_len, trunc)
        assert vol.dtype == np.float32 and vol.shape == (3, 32, 32, 32)
        assert np.array_equal(vol, c["tsdf"])
        # sign of zero included
        assert np.array_equal(np.signbit(vol), np.signbit(c["tsdf"]))


def test_process_frame_end_to_end(forloop_golden):
    # pre/process.py:13-28
    for c in forloop_golden[:3]:
        np.random.seed(int(c["np_random_seed"]))
        pc, vol, max_l, mid = orc.process_frame(_frame(c), int(c["point_num"]))
        assert np.array_equal(pc[:64], c["pc_head"])
        assert np.array_equal(vol, c["tsdf"])
        assert max_l == c["max_l"] and np.array_equal(mid, c["mid_p"])


def test_scalar_loop_equals_vectorised(forloop_golden):
    # the scalar restatement (slow) on one golden frame and at a small R
    c = forloop_golden[3]
    vox_ori, voxel_len, trunc, _, _ = orc.grid_params(c["pmax"], c["pmin"])
    loop = orc.tsdf_cal_loop(c["header"], c["depth"], vox_ori, voxel_len, trunc)
    assert np.array_equal(loop, c["tsdf"])
    for R in (8, 12):
        vox_ori, voxel_len, trunc, _, _ = orc.grid_params(c["pmax"], c["pmin"], R)
        a = orc.tsdf_cal_loop(c["header"], c["depth"], vox_ori, voxel_len, trunc, R)
        b = orc.tsdf_cal(c["header"], c["depth"], vox_ori, voxel_len, trunc, R)
        assert np.array_equal(a, b)


def test_bbox_from_pixels_matches_min_max_kernel(numba_golden):
    # pre/tsdf_numba.py:75-116,140-147 (simulator, Python-2 row index)
    for c in numba_golden:
        pmin, pmax = orc.bbox_from_pixels(c["header"], c["depth"])
        assert np.array_equal(pmin, c["min_p"]) and np.array_equal(pmax, c["max_p"])
        vox_ori, voxel_len, trunc, max_l, mid = orc.grid_params(pmax, pmin)
        assert np.array_equal(vox_ori, c["vox_ori"])
        assert voxel_len == c["voxel_len"] and trunc == c["trunc"]
        assert max_l == c["max_l"] and np.array_equal(mid, c["mid_p"])


def test_numba_layout_is_czyx(numba_golden):
    # pre/tsdf_numba.py:15-72 writes o[c, z, y, x]
    assert sum("tsdf" in c for c in numba_golden) >= 3
    for c in numba_golden:
        if "tsdf" not in c:
            continue
        vol, max_l, mid = orc.dtsdf_from_pixels(_frame(c, "data"), layout=orc.LAYOUT_CZYX)
        assert np.array_equal(vol, c["tsdf"])
        vol0, _, _ = orc.dtsdf_from_pixels(_frame(c, "data"), layout=orc.LAYOUT_CXYZ)
        assert np.array_equal(vol0.transpose(0, 3, 2, 1), c["tsdf"])
        assert max_l == c["max_l"] and np.array_equal(mid, c["mid_p"])


def test_sqrt_truncation_threshold():
    # sqrt_rn(s) > 1  <=>  s > nextafter(1, 2) for binary32 (used by the CUDA kernel
    # to drop the square root; pre/process.py:180-182)
    one = np.float32(1)
    s = one
    vals = [s]
    for _ in range(6):
        s = np.nextafter(s, np.float32(2))
        vals.append(s)
    s = one
    for _ in range(6):
        s = np.nextafter(s, np.float32(0))
        vals.append(s)
    thr = np.nextafter(one, np.float32(2))
    for s in vals:
        assert (np.sqrt(s) > 1) == (s > thr), s
    rng = np.random.default_rng(0)
    s = rng.uniform(0.0, 4.0, size=200000).astype(np.float32)
    assert np.array_equal(np.sqrt(s) > 1, s > thr)
