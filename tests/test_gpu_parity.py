"""Parity of the CUDA path (through the C ABI) with the oracle and with the golden vectors
of the unmodified reference.  Run on the B200 box: python -m pytest tests -m gpu.

Bar: bounding box, grid parameters, voxel/pixel indexing, truncation and sign decisions and
the TSDF values themselves are all BIT-EXACT against the f32 step-wise oracle (tolerance 0,
sign of zero included)."""
import numpy as np
import pytest

from helpers import import_pre, pack_simple, pkg, synth
from oracle import c_oracle
from oracle import dtsdf_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vox():
    return pkg.Voxelizer(device=0)


def _same(a, b):
    """Bit-exact, sign of zero included; NaNs (0/0 when a bounding box degenerates to a point)
    must sit in the same places, their payload is not compared."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b, equal_nan=True):
        return False
    ok = ~np.isnan(a)
    return bool(np.array_equal(np.signbit(a)[ok], np.signbit(b)[ok]))


def test_library_loaded_and_device_present():
    assert pkg.capi.device_count() >= 1
    assert pkg.capi.load_library().dtsdf_abi_version() == 1


# ----------------------------------------------------------------------------- golden vectors
def test_extrema_mode_matches_reference_for_loop(vox, forloop_golden):
    # tsdf_for.tsdf_f / DataProcess.tsdf_f: pre/tsdf_for.py:6-20,44-122
    frames = [{"header": c["header"], "depth": c["depth"]} for c in forloop_golden]
    aux = np.stack([np.concatenate([c["pmax"], c["pmin"]]) for c in forloop_golden]).astype(np.float32)
    vol, max_l, mid_p = vox.voxelize_frames(frames, layout=pkg.LAYOUT_CXYZ, bbox_mode=pkg.BBOX_EXTREMA, aux=aux)
    for i, c in enumerate(forloop_golden):
        assert _same(vol[i], c["tsdf"]), "case %d" % i
        assert max_l[i] == c["max_l"] and np.array_equal(mid_p[i], c["mid_p"])


def test_pixels_mode_matches_reference_numba_kernels(vox, numba_golden):
    # min_max_kernel + tsdf_kernel: pre/tsdf_numba.py:15-116,140-147, layout (c,z,y,x)
    frames = [{"header": c["header"], "data": c["depth"]} for c in numba_golden]
    vol, max_l, mid_p = vox.voxelize_frames(frames, layout=pkg.LAYOUT_CZYX, bbox_mode=pkg.BBOX_PIXELS)
    for i, c in enumerate(numba_golden):
        assert max_l[i] == c["max_l"] and np.array_equal(mid_p[i], c["mid_p"])
        if "tsdf" in c:
            assert _same(vol[i], c["tsdf"]), "case %d" % i


def test_grid_mode_matches_reference_tsdf_cal(vox, forloop_golden):
    # DataProcess.tsdf_cal(vox_ori, voxel_len, trunc): pre/process.py:122-200
    c = forloop_golden[1]
    vox_ori, voxel_len, trunc, _, _ = orc.grid_params(c["pmax"], c["pmin"])
    aux = np.array([[vox_ori[0], vox_ori[1], vox_ori[2], voxel_len, trunc]], dtype=np.float32)
    vol, _, _ = vox.voxelize_frames([{"header": c["header"], "depth": c["depth"]}],
                                    bbox_mode=pkg.BBOX_GRID, aux=aux)
    assert _same(vol[0], c["tsdf"])


# ----------------------------------------------------------------------------- drop-in modules
def test_dataprocess_process_matches_reference(forloop_golden):
    # DataProcess(...).process(): pre/process.py:13-28 with the golden run's np.random seed
    process = import_pre("process")
    for c in forloop_golden:
        frame = {"header": c["header"], "depth": c["depth"]}
        np.random.seed(int(c["np_random_seed"]))
        out = process.DataProcess(frame, np.zeros(63, np.float32), int(c["point_num"]), aug=False).process()
        assert len(out) == 4
        pc, tsdf, max_l, mid_p = out
        assert pc.dtype == np.float64 and pc.shape == (int(c["point_num"]), 3)
        assert np.array_equal(pc[:64], c["pc_head"])
        if "pc" in c:
            assert np.array_equal(pc, c["pc"])
        assert tsdf.shape == (3, 32, 32, 32) and _same(tsdf, c["tsdf"])
        assert max_l == c["max_l"] and np.array_equal(mid_p, c["mid_p"])


def test_dataprocess_methods(forloop_golden):
    process = import_pre("process")
    c = forloop_golden[0]
    frame = {"header": c["header"], "depth": c["depth"]}
    dp = process.DataProcess(frame, None, 6000)
    pts = dp.point_cloud()                                    # pre/process.py:30-68
    assert pts.dtype == np.float64 and pts.shape == (int(c["n_valid"]), 3)
    assert np.array_equal(pts[:64], c["pc_all_head"])
    assert np.array_equal(pts, orc.point_cloud(c["header"], c["depth"]))
    pmax, pmin = dp.max_min_point(c["pc"])                    # pre/process.py:102-120
    assert pmax.dtype == np.float32
    assert np.array_equal(pmax, c["pmax"]) and np.array_equal(pmin, c["pmin"])
    tsdf, max_l, mid_p = dp.tsdf_f(c["pc"])                   # pre/process.py:86-100
    assert _same(tsdf, c["tsdf"]) and max_l == c["max_l"]
    vox_ori, voxel_len, trunc, _, _ = orc.grid_params(c["pmax"], c["pmin"])
    assert _same(dp.tsdf_cal(vox_ori, voxel_len, trunc), c["tsdf"])   # pre/process.py:122
    # cut_version signature and aliases (cut_version/pre/process.py:6,52)
    np.random.seed(int(c["np_random_seed"]))
    pc = process.DataProcess(frame, 6000, False).get_points()
    assert np.array_equal(pc, c["pc"])
    with pytest.raises(NotImplementedError):
        process.DataProcess(frame, None, 6000, aug=True).process()


def test_dataprocess_instances_interleaved(forloop_golden):
    """Two DataProcess objects (and a tsdf_for call) share the process-wide device context: calls of one
    between point_cloud() and set_length() of the other, and in-place edits of the arrays handed back,
    must not leak into a result - the reference always works on the array it is handed
    (pre/process.py:70-84,102-120)."""
    process, tsdf_for = import_pre("process"), import_pre("tsdf_for")
    a, b = forloop_golden[0], forloop_golden[3]
    fa = {"header": a["header"], "depth": a["depth"]}
    fb = {"header": b["header"], "depth": b["depth"]}
    dpa, dpb = process.DataProcess(fa, None, int(a["point_num"])), process.DataProcess(fb, None, int(b["point_num"]))
    hpa = dpa.point_cloud()
    hpb = dpb.point_cloud()                                   # overwrites the cloud resident in the context
    tsdf_for.max_min_point(b["pc"])                           # and so does this
    np.random.seed(int(a["np_random_seed"]))
    pca = dpa.set_length(hpa)
    assert np.array_equal(pca[:64], a["pc_head"])
    if "pc" in a:
        assert np.array_equal(pca, a["pc"])
    np.random.seed(int(b["np_random_seed"]))
    pcb = dpb.set_length(hpb)
    ta = dpa.tsdf_f(pca)                                      # after another instance's set_length
    assert _same(ta[0], a["tsdf"]) and ta[1] == a["max_l"]
    assert _same(dpb.tsdf_f(pcb)[0], b["tsdf"])
    # in-place edits count: the gather and the extrema follow the array's current content
    hp2 = hpa.copy()
    hp2[:, 0] += 5.0
    np.random.seed(int(a["np_random_seed"]))
    idx = np.random.randint(0, hp2.shape[0], size=dpa.point_num) if hp2.shape[0] >= dpa.point_num else None
    hpa[:, 0] += 5.0
    np.random.seed(int(a["np_random_seed"]))
    moved = dpa.set_length(hpa)
    if idx is not None:
        assert np.array_equal(moved, hp2[idx])
    pca[:, 2] -= 3.0
    pmax, pmin = dpa.max_min_point(pca)
    z = pca[:, 2][pca[:, 2] != 0]
    assert pmax[2] == np.float32(z.max()) and pmin[2] == np.float32(z.min())


def test_process_host_one_call_equals_per_frame_numpy_draws():
    """dtsdf_process_host (index vectors drawn by the library from the caller's MT19937 state, chunk by chunk,
    pipelined) against dtsdf_process_begin_host + one np.random.randint call per frame +
    dtsdf_process_finish_host: the same clouds, volumes, max_l, mid_p and the same np.random state afterwards.
    Packed and plain routes, several chunks, frames with fewer points than point_num."""
    frames = synth.make_frames(range(1200, 1200 + 150))
    small = synth.make_frame(3, box=(150, 100, 190, 140))              # 1,600 px: fewer points than 6000
    frames[5] = small
    frames[77] = small
    v = pkg.Voxelizer(device=0)
    for chunk_mb, n, want_points, packed_route in ((64, 150, True, 1), (2, 150, False, 1), (1, 150, True, 1),
                                                   (64, 30, True, 1), (1, 60, True, 0)):
        v.set_option(pkg.capi.OPT_CHUNK_MB, chunk_mb)
        v.set_option(pkg.capi.OPT_PACKED_D2H, packed_route)
        np.random.seed(99)
        a = v.process_frames_two_calls(frames[:n], 6000, want_points=want_points)
        tail_a = np.random.randint(0, 1 << 30, size=4)
        np.random.seed(99)
        b = v.process_frames(frames[:n], 6000, want_points=want_points)
        tail_b = np.random.randint(0, 1 << 30, size=4)
        assert np.array_equal(tail_a, tail_b)
        assert (a[0] is None) == (b[0] is None) == (not want_points)
        if want_points:
            assert np.array_equal(a[0], b[0])
        assert _same(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    # a crop beyond what dtsdf_sample_cloud_batch takes (a full 320 x 240 frame): the call builds the full clouds
    # up front and gathers from them, same results
    big = frames[:5] + [synth.make_frame(1, box=(0, 0, 320, 240))] + frames[5:9]
    v.set_option(pkg.capi.OPT_CHUNK_MB, 1)
    v.set_option(pkg.capi.OPT_PACKED_D2H, 1)
    np.random.seed(7)
    a = v.process_frames_two_calls(big, 6000, want_points=True)
    np.random.seed(7)
    b = v.process_frames(big, 6000, want_points=True)
    assert np.array_equal(a[0], b[0]) and _same(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    # a frame without a single point: the reference's np.random.randint(0, 0) raises ValueError
    empty = synth.make_frame(4)
    empty["depth"] = np.zeros_like(empty["depth"])
    with pytest.raises(ValueError):
        v.process_frames(frames[:3] + [empty] + frames[3:6], 6000)
    np.random.seed(1)
    out = v.process_frames(frames[:4], 6000)                              # the context works afterwards
    np.random.seed(1)
    assert _same(out[1], v.process_frames_two_calls(frames[:4], 6000)[1])
    # the same failure deep inside a pipelined call (two frames per chunk, the index scan two chunks ahead of the
    # submission, crops uploaded chunk by chunk): the caller's stream ends where the reference's would - advanced by
    # the draws of the frames before the empty one, nothing more
    v.set_option(pkg.capi.OPT_CHUNK_MB, 1)
    v.set_option(pkg.capi.OPT_PACKED_D2H, 1)
    batch = frames[:23] + [empty] + frames[23:40]
    np.random.seed(123)
    with pytest.raises(ValueError):
        v.process_frames(batch, 6000, want_points=False)
    mine = np.random.get_state()
    np.random.seed(123)
    for fr in batch[:23]:
        ps = int(np.count_nonzero(fr["depth"]))
        np.random.randint(0, ps, size=6000 - ps if ps < 6000 else 6000)
    theirs = np.random.get_state()
    assert np.array_equal(mine[1], theirs[1]) and mine[2] == theirs[2]
    v.close()


def _sampled_case_frames():
    """Hands, frames that look nothing like a hand (boxes partly outside the image, holes, values straddling
    |d| = 1, negative depths), NaN and huge depths, mixed signs in the centre column / row (products of either
    zero), a full 320x240 frame (not staged in shared memory), 1-pixel crops and a 1-pixel-wide one."""
    rng = np.random.default_rng(2024)
    frames = synth.make_frames(range(1500, 1512)) + _random_frames(rng, 40)
    f = synth.make_frame(7)
    d = f["depth"].copy()
    d[rng.integers(0, d.size, 30)] = np.nan
    d[rng.integers(0, d.size, 30)] = 3e12
    d[rng.integers(0, d.size, 30)] = -450.0
    d[rng.integers(0, d.size, 30)] = 0.25                      # a point of the cloud, no depth for the voxel loop
    frames.append({"header": f["header"], "depth": d})
    yy, xx = np.mgrid[0:40, 0:60].astype(np.float32)           # box around the principal point, depth changing sign
    d = np.where((xx + yy) % 3 == 0, -300.0 - xx, 280.0 + yy).astype(np.float32)
    frames.append({"header": np.array([320, 240, 130, 100, 190, 140], np.int32), "depth": d.reshape(-1)})
    frames.append(synth.make_frame(1, box=(0, 0, 320, 240)))
    frames.append({"header": np.array([320, 240, 17, 33, 18, 34], np.int32), "depth": np.array([412.5], np.float32)})
    frames.append({"header": np.array([320, 240, 200, 10, 201, 110], np.int32),
                   "depth": rng.uniform(300, 500, size=100).astype(np.float32)})
    frames.append({"header": np.array([321, 239, 100, 50, 160, 120], np.int32),      # odd image size: half-integer u, v
                   "depth": rng.uniform(200, 600, size=60 * 70).astype(np.float32)})
    return frames


@pytest.mark.parametrize("R,layout", [(32, 0), (32, 1), (48, 0), (64, 1)])
def test_sampled_voxeliser_equals_the_cloud_path(R, layout):
    """DTSDF_BBOX_SAMPLED (dtsdf_count_valid_batch + dtsdf_voxelize_sampled_batch: ranks, sample bitmap and
    extrema inside the voxeliser launch, no cloud in memory) against the three-kernel form that materialises
    float64 clouds (dtsdf_process_begin_host / _finish_host) for the same index vectors, bit for bit; and,
    on frames without NaN depths, against the NumPy restatement of point_cloud / set_length / max_min_point
    (pre/process.py:30-84,102-120) followed by the C oracle's voxel loop."""
    import torch
    v = pkg.Voxelizer(device=0, voxel_res=R)
    frames = _sampled_case_frames()
    n_valid = v.process_begin(frames)
    for f, nv in zip(frames, n_valid):
        d = np.asarray(f["depth"])
        assert nv == int(np.count_nonzero(d != 0))             # NaN != 0: a row any() keeps (pre/process.py:63)
    frames = [f for f, nv in zip(frames, n_valid) if nv > 0]
    n_valid = v.process_begin(frames)
    n = len(frames)
    rng = np.random.default_rng(R + layout)
    for point_num in (6000, 37):
        idx = np.empty((n, point_num), np.int32)
        for i, nv in enumerate(n_valid):
            if nv < point_num:                                  # pre/process.py:74-77
                idx[i] = np.arange(point_num)
                idx[i, nv:] = rng.integers(0, nv, size=point_num - nv)
            else:
                idx[i] = rng.integers(0, nv, size=point_num)
        pcs, want_vol, want_ml, want_mp = v.process_finish(idx, point_num, True, layout, R)
        headers, offsets, packed, max_px = pkg.pack_frames(frames)
        dev = torch.device("cuda:0")
        args = (torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev))
        assert np.array_equal(pkg.count_valid_device(*args).cpu().numpy(), n_valid)
        cfg = pkg.capi.default_config(R, layout, pkg.BBOX_PIXELS)
        for hint in (max_px, 64):                               # 64: nothing fits the staging buffer
            vol, ml, mp = pkg.voxelize_sampled_device(*args, torch.from_numpy(idx).to(dev), hint, cfg)
            torch.cuda.synchronize()
            assert _same(ml.cpu().numpy(), want_ml) and _same(mp.cpu().numpy(), want_mp)
            got = vol.cpu().numpy()
            for i in range(n):
                assert _same(got[i], want_vol[i]), "frame %d hint %d" % (i, hint)
        # the restatement of the reference's own steps
        aux = np.zeros((n, 6), np.float32)
        clean = []
        for i, f in enumerate(frames):
            d = np.asarray(f["depth"])
            if np.isnan(d).any():
                continue
            pc = orc.point_cloud(f["header"], d)[idx[i]]
            assert np.array_equal(pc, pcs[i])
            pmax, pmin = orc.max_min_point(pc)
            aux[i, :3], aux[i, 3:] = pmax, pmin
            clean.append(i)
        cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed, R=R, layout=layout, bbox_mode=1, aux=aux)
        for i in clean:
            assert _same(got[i], cv[i]) and _same(want_ml[i], cml[i]) and _same(want_mp[i], cmp_[i]), "oracle frame %d" % i
    v.close()


def test_sampled_cloud_in_one_launch_equals_the_two_kernels():
    """dtsdf_sample_cloud_batch (hand_points[rand_idx] straight from the depth map: word masks, rank search,
    back-projection of the sampled pixels only) against dtsdf_point_cloud_batch + dtsdf_resample_extrema_batch, bit
    for bit: sampled clouds, extrema, valid-pixel counts; hands, random boxes, NaN / huge / negative / sub-unit
    depths, 1-pixel crops, an odd image size; 6000 and 37 samples; a crop beyond its limit is refused."""
    import ctypes
    import torch
    lib = pkg.capi.load_library()
    dev = torch.device("cuda:0")
    frames = [f for f in _sampled_case_frames() if np.asarray(f["depth"]).size <= 22528 and np.count_nonzero(f["depth"])]
    assert len(frames) > 50
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    n = len(frames)
    d_packed, d_off, d_hdr = torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev)
    sizes = (headers[:, 4] - headers[:, 2]).astype(np.int64) * (headers[:, 5] - headers[:, 3])
    poff = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    d_poff = torch.from_numpy(poff).to(dev)
    d_points = torch.zeros((int(sizes.sum()), 3), dtype=torch.float64, device=dev)
    d_nvalid = torch.zeros((n,), dtype=torch.int32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    focal = ctypes.c_double(241.42)
    pkg.capi.check(lib.dtsdf_point_cloud_batch(n, focal, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(),
                                               d_points.data_ptr(), d_poff.data_ptr(), d_nvalid.data_ptr(), stream))
    n_valid = d_nvalid.cpu().numpy()
    rng = np.random.default_rng(99)
    for point_num in (6000, 37):
        idx = np.stack([rng.integers(0, int(v), size=point_num) for v in n_valid]).astype(np.int32)
        idx[:, 0] = 0
        idx[:, 1] = n_valid - 1                                  # first and last point of every frame
        d_idx = torch.from_numpy(idx).to(dev)
        want_pc = torch.zeros((n, point_num, 3), dtype=torch.float64, device=dev)
        want_ext = torch.zeros((n, 6), dtype=torch.float32, device=dev)
        pkg.capi.check(lib.dtsdf_resample_extrema_batch(n, point_num, d_points.data_ptr(), d_poff.data_ptr(), d_idx.data_ptr(),
                                                        want_pc.data_ptr(), want_ext.data_ptr(), stream))
        got_pc = torch.full((n, point_num, 3), 7.0, dtype=torch.float64, device=dev)
        got_ext = torch.full((n, 6), 7.0, dtype=torch.float32, device=dev)
        got_nv = torch.zeros((n,), dtype=torch.int32, device=dev)
        pkg.capi.check(lib.dtsdf_sample_cloud_batch(n, focal, point_num, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(),
                                                    d_idx.data_ptr(), int(max_px), got_pc.data_ptr(), got_ext.data_ptr(),
                                                    got_nv.data_ptr(), stream))
        torch.cuda.synchronize()
        assert np.array_equal(got_nv.cpu().numpy(), n_valid)
        assert torch.equal(got_pc.view(torch.int64), want_pc.view(torch.int64))
        assert torch.equal(got_ext.view(torch.int32), want_ext.view(torch.int32))
        # without a cloud output, and with ranks that are not points of the frame
        idx[::2, 5] = n_valid[::2]
        idx[1::2, 6] = -3
        d_idx = torch.from_numpy(idx).to(dev)
        pkg.capi.check(lib.dtsdf_sample_cloud_batch(n, focal, point_num, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(),
                                                    d_idx.data_ptr(), int(max_px), got_pc.data_ptr(), got_ext.data_ptr(), None, stream))
        torch.cuda.synchronize()
        got = got_pc.cpu().numpy()
        assert not got[::2, 5].any() and not got[1::2, 6].any()
        keep = np.ones((n, point_num), bool)
        keep[::2, 5] = False
        keep[1::2, 6] = False
        assert np.array_equal(got.view(np.int64)[keep], want_pc.cpu().numpy().view(np.int64)[keep])
    rc = lib.dtsdf_sample_cloud_batch(n, focal, 37, d_packed.data_ptr(), d_off.data_ptr(), d_hdr.data_ptr(), d_idx.data_ptr(),
                                      76800, got_pc.data_ptr(), got_ext.data_ptr(), None, stream)
    assert rc == pkg.capi.ERR_UNSUPPORTED


def test_kernels_stay_inside_their_buffers():
    """compute-sanitizer is closed on this pool (profiles/r02_sanitizer_refused.txt): every output of every device
    entry point is carved out of a larger allocation filled with a canary pattern, 4 KB on both sides; after the
    launches over edge shapes (1-pixel crops, a full frame that is not staged, crops partly outside the image, empty
    depth) the canaries must be untouched and the results must equal those written into ordinary tensors."""
    import importlib
    import torch
    amod = importlib.import_module(pkg.__name__ + ".augment")
    cmod = importlib.import_module(pkg.__name__ + ".codec")
    dev = torch.device("cuda:0")
    CANARY, PAD = 0x7FC0DEAD, 1024                              # int32 words
    frames = _sampled_case_frames()
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    n = len(frames)
    args = (torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev))
    guards = []

    def guarded(shape, dtype=torch.float32):
        numel = int(np.prod(shape))
        assert dtype in (torch.float32, torch.int32)
        raw = torch.full((numel + 2 * PAD,), CANARY, dtype=torch.int32, device=dev)
        guards.append((raw, numel))
        return raw[PAD:PAD + numel].view(dtype).view(*shape)

    def check_guards():
        torch.cuda.synchronize()
        for raw, numel in guards:
            assert bool((raw[:PAD] == CANARY).all()) and bool((raw[PAD + numel:] == CANARY).all())
        guards.clear()

    n_valid = pkg.count_valid_device(*args, out=guarded((n,), torch.int32))
    check_guards()
    nv = n_valid.cpu().numpy()
    keep = nv > 0
    rng = np.random.default_rng(5)
    idx = np.stack([rng.integers(0, max(int(v), 1), size=6000) for v in nv]).astype(np.int32)
    idx[::3, :7] = [-1, -2 ** 31, 2 ** 31 - 1, 76800, 10 ** 6, 123456789, -5]      # out of range: must select nothing
    d_idx = torch.from_numpy(idx).to(dev)
    for R, layout in ((32, 0), (48, 1), (64, 0)):
        cfg = pkg.capi.default_config(R, layout, pkg.BBOX_PIXELS)
        for hint in (max_px, 16):
            out = (guarded((n, 3, R, R, R)), guarded((n,)), guarded((n, 3)))
            pkg.voxelize_device(*args, hint, cfg, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2])
            want = pkg.voxelize_device(*args, hint, cfg)
            torch.cuda.synchronize()
            assert all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(out, want))
            check_guards()
            out = (guarded((n, 3, R, R, R)), guarded((n,)), guarded((n, 3)))
            pkg.voxelize_sampled_device(*args, d_idx, hint, cfg, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2])
            want = pkg.voxelize_sampled_device(*args, d_idx, hint, cfg)
            torch.cuda.synchronize()
            sel = torch.from_numpy(keep).to(dev)
            assert torch.equal(out[0][sel].view(torch.int32), want[0][sel].view(torch.int32))
            check_guards()
            aug = torch.from_numpy(amod.make_augmentations(n, np.random.default_rng(R), centers=np.tile([0, 0, -380.0], (n, 1)))).to(dev)
            out = (guarded((n, 3, R, R, R)), guarded((n,)), guarded((n, 3)))
            amod.voxelize_aug_device(*args, aug, hint, cfg, out_vol=out[0], out_max_l=out[1], out_mid_p=out[2])
            check_guards()
    # the volume codec: packed arrays carved from a guarded byte buffer, expansion into a guarded volume
    vol, _, _ = pkg.voxelize_device(*args, max_px)
    nbytes = cmod.PackedVolumes.storage_bytes(n, 32)
    raw = torch.full((nbytes // 4 + 2 * PAD,), CANARY, dtype=torch.int32, device=dev)
    store = raw[PAD:PAD + nbytes // 4].view(torch.uint8)
    pk = cmod.PackedVolumes(n, 32, dev, storage=store)
    cmod.pack_volumes_device(vol, pk)
    back = guarded((n, 3, 32, 32, 32))
    cmod.unpack_volumes_device(pk, out=back)
    torch.cuda.synchronize()
    assert bool((raw[:PAD] == CANARY).all()) and bool((raw[PAD + nbytes // 4:] == CANARY).all())
    assert torch.equal(back.view(torch.int32), vol.view(torch.int32))
    check_guards()


def test_tsdf_for_module(forloop_golden):
    tsdf_for = import_pre("tsdf_for")
    c = forloop_golden[3]
    frame = {"header": c["header"], "depth": c["depth"]}
    tsdf, max_l, mid_p = tsdf_for.tsdf_f(frame, c["pc"])      # pre/tsdf_for.py:6-20
    assert _same(tsdf, c["tsdf"]) and max_l == c["max_l"] and np.array_equal(mid_p, c["mid_p"])
    pmax, pmin = tsdf_for.max_min_point(c["pc"])
    assert np.array_equal(pmax, c["pmax"]) and np.array_equal(pmin, c["pmin"])


def test_tsdf_numba_module(numba_golden):
    tsdf_numba = import_pre("tsdf_numba")
    assert (tsdf_numba.FOCAL, tsdf_numba.CENTER_X, tsdf_numba.CENTER_Y, tsdf_numba.VOXEL_RES) == (241.42, 160, 120, 32)
    c = numba_golden[0]
    tsdf, max_l, mid_p = tsdf_numba.cal_tsdf_cuda({"header": c["header"], "data": c["depth"]})
    assert tsdf.dtype == np.float32 and _same(tsdf, c["tsdf"])
    assert max_l == c["max_l"] and np.array_equal(mid_p, c["mid_p"])
    # error convention: None (after printing) when there is no valid pixel, pre/tsdf_numba.py:162-171
    empty = {"header": c["header"], "data": np.zeros_like(c["depth"])}
    assert tsdf_numba.cal_tsdf_cuda(empty) is None
    # anything else propagates as an exception
    with pytest.raises(KeyError):
        tsdf_numba.cal_tsdf_cuda({"header": c["header"], "depth": c["depth"]})


# ----------------------------------------------------------------------------- oracle sweeps
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("R", [32, 48, 64])
def test_ragged_batch_matches_oracle(vox, R, layout):
    frames = synth.make_frames(range(200, 232))
    vol, max_l, mid_p = vox.voxelize_frames(frames, layout=layout, voxel_res=R)
    for i, fr in enumerate(frames[:8]):                       # NumPy oracle on a few
        v, ml, mp = orc.dtsdf_from_pixels(fr, R=R, layout=layout)
        assert _same(vol[i], v) and max_l[i] == ml and np.array_equal(mid_p[i], mp)
    headers, offsets, packed = pack_simple(frames)            # C oracle on all
    cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed, R=R, layout=layout)
    assert _same(vol, cv) and np.array_equal(max_l, cml) and np.array_equal(mid_p, cmp_)


def test_batch_larger_than_grid(vox):
    # more frames than resident CTAs: every CTA loops over several frames
    frames = synth.make_frames(range(1000, 1000 + 1500))
    headers, offsets, packed = pack_simple(frames)
    vol, max_l, mid_p = vox.voxelize_packed(headers, offsets, packed)
    cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed)
    assert _same(vol, cv) and np.array_equal(max_l, cml) and np.array_equal(mid_p, cmp_)


def test_edge_crops(vox):
    frames = [
        synth.make_frame(1, box=(0, 0, 320, 240)),             # full image: larger than shared memory, direct path
        synth.make_frame(2, box=(40, 20, 260, 230)),           # 220x210 = 185 KB: staged, one CTA per SM
        synth.make_frame(3, box=(159, 119, 160, 120)),         # 1x1 crop
        synth.make_frame(4, box=(100, 100, 103, 101)),         # 3x1 crop (padded to 4 floats)
        synth.make_frame(5, box=(10, 10, 11, 200)),            # 1 px wide
        synth.make_frame(6, box=(0, 0, 96, 80)),               # touching the image corner
    ]
    z = synth.make_frame(7)
    z["depth"] = np.zeros_like(z["depth"])                     # no valid pixel at all
    frames.append(z)
    s = synth.make_frame(8)
    s["depth"] = np.where(s["depth"] > 0, np.float32(0.5), np.float32(0))   # |d| < 1 everywhere
    frames.append(s)
    n = synth.make_frame(9)
    n["depth"] = -n["depth"]                                   # negative depths are valid (abs() test)
    frames.append(n)
    headers, offsets, packed = pack_simple(frames)
    for layout in (0, 1):
        vol, max_l, mid_p = vox.voxelize_packed(headers, offsets, packed, layout=layout)
        cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed, layout=layout)
        for i in range(len(frames)):
            assert _same(vol[i], cv[i]), "frame %d layout %d" % (i, layout)
        assert np.array_equal(max_l, cml) and np.array_equal(mid_p, cmp_)


def test_nan_and_huge_depths(vox):
    """Pixels holding NaN are not 'no depth' in the reference (abs(d) < 1 is false for NaN,
    pre/process.py:167): the voxels that project onto them come out NaN.  Such a crop leaves
    the sanitised staged path; depths beyond 2^40 leave the fixed-divisor scan."""
    a = synth.make_frame(11)
    a["depth"] = a["depth"].copy()
    valid = np.flatnonzero(a["depth"] > 0)
    a["depth"][valid[::37]] = np.nan
    b = synth.make_frame(12)
    b["depth"] = b["depth"].copy()
    valid = np.flatnonzero(b["depth"] > 0)
    b["depth"][valid[5]] = np.float32(3e12)
    b["depth"][valid[9]] = np.float32(-2e13)
    c = synth.make_frame(13)                                   # an ordinary neighbour in the same batch
    headers, offsets, packed = pack_simple([a, b, c, a])
    for layout in (0, 1):
        vol, max_l, mid_p = vox.voxelize_packed(headers, offsets, packed, layout=layout)
        cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed, layout=layout)
        assert np.isnan(cv[0]).any() and not np.isnan(cv[2]).any()
        for i in range(4):
            assert _same(vol[i], cv[i]), "frame %d layout %d" % (i, layout)
        assert np.array_equal(max_l, cml) and np.array_equal(mid_p, cmp_)


def test_empty_batch_and_bad_arguments(vox):
    vol, max_l, mid_p = vox.voxelize_frames([])
    assert vol.shape == (0, 3, 32, 32, 32) and max_l.shape == (0,)
    fr = synth.make_frame(0)
    with pytest.raises(pkg.DtsdfError) as e:
        vox.voxelize_frames([fr], voxel_res=40)
    assert e.value.code == pkg.capi.ERR_UNSUPPORTED
    bad = {"header": np.array([320, 240, 100, 60, 100, 190], np.int32), "depth": np.zeros(0, np.float32)}
    with pytest.raises(pkg.DtsdfError) as e:
        vox.voxelize_frames([bad])
    assert e.value.code == pkg.capi.ERR_BAD_HEADER
    with pytest.raises(ValueError):
        vox.voxelize_frames([{"header": fr["header"], "depth": fr["depth"][:-1]}])


def test_packed_and_plain_host_transfer_agree():
    """dtsdf_voxelize_host brings volumes back packed (2-bit codes + literals, expanded by host
    threads) or as plain copies (DTSDF_OPT_PACKED_D2H): the same bits either way, several chunks,
    every resolution and layout, odd thread counts, an unaligned destination."""
    frames = synth.make_frames(range(500, 500 + 400))          # 400 x 393 KB = 2.4 chunks at R = 32
    z = synth.make_frame(7)
    z["depth"] = np.zeros_like(z["depth"])
    frames[13] = z                                              # an all-zero volume: no literals at all
    headers, offsets, packed = pack_simple(frames)
    v = pkg.Voxelizer(device=0)
    for R, layout, nthreads, nfr in ((32, 0, 0, 400), (32, 1, 3, 400), (48, 0, 5, 60), (64, 1, 1, 30)):
        hdr, off = headers[:nfr], offsets[:nfr]
        v.set_option(pkg.capi.OPT_PACKED_D2H, 0)
        plain = v.voxelize_packed(hdr, off, packed, layout=layout, voxel_res=R)
        v.set_option(pkg.capi.OPT_PACKED_D2H, 1)
        v.set_option(pkg.capi.OPT_HOST_THREADS, nthreads)
        got = v.voxelize_packed(hdr, off, packed, layout=layout, voxel_res=R)
        for a, b in zip(plain, got):
            assert _same(a, b)
        raw = np.empty(nfr * 3 * R ** 3 + 1, dtype=np.float32)            # destination only 4-byte aligned
        dst = raw[1:].reshape(nfr, 3, R, R, R)
        v.voxelize_packed(hdr, off, packed, layout=layout, voxel_res=R, out_vol=dst)
        assert _same(dst, plain[0])
    cv, _, _, _ = c_oracle.dtsdf_batch(headers, offsets, packed)
    v.set_option(pkg.capi.OPT_HOST_THREADS, 0)
    assert _same(v.voxelize_packed(headers, offsets, packed)[0], cv)
    with pytest.raises(pkg.DtsdfError):
        v.set_option(99, 1)
    v.close()


def test_host_pipeline_many_small_chunks_and_destination_alignments():
    """The host pipeline hands chunks to the expansion threads from a finalizer thread and
    reuses its NSLOT chunk buffers: 1 MB chunks (2 frames each, ~150 chunks through 4 slots),
    repeated calls on one context, destinations at every 16-byte phase of a 64-byte line (the
    AVX-512 expansion peels code bytes up to the line boundary) - always the device run's bits."""
    frames = synth.make_frames(range(900, 900 + 301))
    headers, offsets, packed = pack_simple(frames)
    v = pkg.Voxelizer(device=0)
    want = c_oracle.dtsdf_batch(headers, offsets, packed)[0]
    v.set_option(pkg.capi.OPT_CHUNK_MB, 1)
    n, vf = len(frames), 3 * 32 ** 3
    raw = np.empty(n * vf + 64, dtype=np.float32)
    base = (-raw.ctypes.data // 4) % 16                     # index of a 64-byte aligned float
    for nthreads, phase in ((0, 0), (1, 4), (7, 8), (16, 12), (3, 0)):
        v.set_option(pkg.capi.OPT_HOST_THREADS, nthreads)
        dst = raw[base + phase: base + phase + n * vf].reshape(n, 3, 32, 32, 32)
        assert dst.ctypes.data % 64 == 4 * phase
        dst[...] = np.float32(7.0)
        _, max_l, mid_p = v.voxelize_packed(headers, offsets, packed, out_vol=dst)
        assert _same(dst, want)
    v.set_option(pkg.capi.OPT_CHUNK_MB, 64)
    assert _same(v.voxelize_packed(headers, offsets, packed)[0], want)
    v.close()


def test_packed_host_result_expands_to_the_dense_result():
    """dtsdf_voxelize_host_packed + dtsdf_expand_frames: the volumes stay packed in host memory and any frame
    range expands to exactly what dtsdf_voxelize_host returns; small batches, several chunks, every
    resolution, both literal transfer routes (copy engine / push kernel)."""
    frames = synth.make_frames(range(4100, 4100 + 260))
    headers, offsets, packed = pack_simple(frames)
    v = pkg.Voxelizer(device=0)
    for R, n, chunk_mb, push in ((32, 260, 8, 0), (32, 260, 64, 1), (32, 3, 64, 0), (48, 40, 16, 0), (64, 20, 64, 0)):
        v.set_option(pkg.capi.OPT_CHUNK_MB, chunk_mb)
        v.set_option(pkg.capi.OPT_LITERAL_PUSH, push)
        want, wml, wmp = v.voxelize_packed(headers[:n], offsets[:n], packed, voxel_res=R)
        lazy, ml, mp = v.voxelize_packed_lazy(headers[:n], offsets[:n], packed, voxel_res=R)
        assert len(lazy) == n and lazy.shape == want.shape and lazy.nbytes < want.nbytes / 4
        assert np.array_equal(ml, wml) and np.array_equal(mp, wmp)
        assert _same(lazy.expand(), want)
        assert _same(lazy.expand(threads=3), want)
        a, b = n // 3, n // 3 + min(7, n - n // 3)
        assert _same(lazy[a:b], want[a:b]) and _same(lazy[n - 1], want[n - 1]) and _same(lazy[-1], want[-1])
        assert _same(lazy[np.array([0, n - 1, n // 2])], want[[0, n - 1, n // 2]])
        with pytest.raises(pkg.DtsdfError):
            lazy.expand(n - 1, 2)
        lazy.close()
    # the storage of closed results is kept for later ones (no fresh pages per call): results that are alive at
    # the same time never share it, and a recycled block holds nothing of its previous owner
    v.set_option(pkg.capi.OPT_CHUNK_MB, 8)
    want = v.voxelize_packed(headers, offsets, packed)[0]
    first = v.voxelize_packed_lazy(headers[:200], offsets[:200], packed)[0]
    second = v.voxelize_packed_lazy(headers[60:260], offsets[60:260], packed)[0]
    assert _same(first.expand(), want[:200]) and _same(second.expand(), want[60:260])
    first.close()
    third = v.voxelize_packed_lazy(headers[100:130], offsets[100:130], packed)[0]       # takes blocks `first` gave back
    assert _same(third.expand(), want[100:130]) and _same(second.expand(), want[60:260])
    second.close()
    third.close()
    v.set_option(pkg.capi.OPT_LITERAL_PUSH, 1)                    # the push route of the dense result, too
    v.set_option(pkg.capi.OPT_CHUNK_MB, 8)
    assert _same(v.voxelize_packed(headers, offsets, packed)[0], c_oracle.dtsdf_batch(headers, offsets, packed)[0])
    v.close()


def test_host_pipeline_failure_leaves_no_writer_behind():
    """A failure in the middle of the chunk loop (injected: DTSDF_OPT_TEST_FAIL_CHUNK) with earlier chunks in
    flight: the call returns the error only after every copy and expansion thread is done or cancelled -
    nothing writes into the caller's buffers after the return - and the context works afterwards.  Packed
    and plain routes, pageable max_l / mid_p."""
    import time
    frames = synth.make_frames(range(7000, 7000 + 240))
    headers, offsets, packed = pack_simple(frames)
    want, wml, wmp, _ = c_oracle.dtsdf_batch(headers, offsets, packed)
    v = pkg.Voxelizer(device=0)
    v.set_option(pkg.capi.OPT_CHUNK_MB, 4)                       # 10 frames per chunk, 24 chunks
    out = np.empty_like(want)
    ml, mp = np.empty(len(frames), np.float32), np.empty((len(frames), 3), np.float32)
    for packed_route in (1, 0):
        v.set_option(pkg.capi.OPT_PACKED_D2H, packed_route)
        for fail_at in (0, 3, 7, 23):
            v.set_option(pkg.capi.OPT_TEST_FAIL_CHUNK, fail_at)
            with pytest.raises(pkg.DtsdfError) as ei:
                v.voxelize_packed(headers, offsets, packed, out_vol=out, out_max_l=ml, out_mid_p=mp)
            assert ei.value.code == pkg.capi.ERR_CUDA
            out[...] = np.float32(-7.0)
            ml[...] = -7.0
            mp[...] = -7.0
            time.sleep(0.05)
            assert (out == np.float32(-7.0)).all() and (ml == -7.0).all() and (mp == -7.0).all()
        v.set_option(pkg.capi.OPT_TEST_FAIL_CHUNK, -1)
        v.voxelize_packed(headers, offsets, packed, out_vol=out, out_max_l=ml, out_mid_p=mp)
        assert _same(out, want) and np.array_equal(ml, wml) and np.array_equal(mp, wmp)
    v.close()


def test_host_entry_points_restore_the_current_device():
    """A process that drives several GPUs: the library's host entry points run on the context's device
    and put the caller's current device back."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    frames = synth.make_frames(range(10, 14))
    torch.cuda.set_device(0)
    v1 = pkg.Voxelizer(device=1)
    assert torch.cuda.current_device() == 0
    got = v1.voxelize_frames(frames)[0]
    v1.point_cloud(frames[0])
    assert torch.cuda.current_device() == 0
    x = torch.zeros(4, device="cuda")
    assert x.device.index == 0
    assert _same(got, c_oracle.dtsdf_batch(*pack_simple(frames))[0])
    # every kernel of the process() path on the second device as well (kernel attributes such as the opt-in shared
    # memory limit are per device), same results as on the first
    v0 = pkg.Voxelizer(device=0)
    for want_points in (True, False):
        np.random.seed(3)
        a = v0.process_frames(frames, 6000, want_points=want_points)
        np.random.seed(3)
        b = v1.process_frames(frames, 6000, want_points=want_points)
        assert torch.cuda.current_device() == 0
        assert (a[0] is None and b[0] is None) or np.array_equal(a[0], b[0])
        assert _same(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    v0.close()
    v1.close()
    assert torch.cuda.current_device() == 0


def test_host_pipeline_random_configurations():
    """Random batch sizes (plain and packed routes), chunk sizes and expansion thread counts on one
    context, back to back: the host entry point always returns the oracle's bits."""
    frames = synth.make_frames(range(3000, 3000 + 400))
    headers, offsets, packed = pack_simple(frames)
    want, wml, wmp, _ = c_oracle.dtsdf_batch(headers, offsets, packed)
    v = pkg.Voxelizer(device=0)
    rng = np.random.default_rng(5)
    out = np.empty_like(want)
    for it in range(60):
        mb = int(rng.choice([1, 2, 3, 5, 8, 16, 33, 64, 200]))
        nt = int(rng.integers(0, 17))
        n = int(rng.integers(1, 401))
        v.set_option(pkg.capi.OPT_CHUNK_MB, mb)
        v.set_option(pkg.capi.OPT_HOST_THREADS, nt)
        out[:n] = 5.0
        _, ml, mp = v.voxelize_packed(headers[:n], offsets[:n], packed, out_vol=out[:n])
        assert _same(out[:n], want[:n]), "call %d: chunk %d MB, %d threads, %d frames" % (it, mb, nt, n)
        assert np.array_equal(ml, wml[:n]) and np.array_equal(mp, wmp[:n])
    v.close()


def test_joint_nor_module(tmp_path, monkeypatch):
    # pre/joint_nor.py:8-18 (float64 there): normalize() and the result-tree walk of read_joint()
    jn = import_pre("joint_nor")
    rng = np.random.default_rng(11)
    n = 37
    gt = (rng.normal(size=(n, 21, 3)) * 60 + np.array([10, -20, -400])).astype(np.float32)
    max_l = rng.uniform(120, 260, size=n)
    mid_p = rng.normal(size=(n, 3)) * 30 + np.array([0, 0, -400.0])
    want = np.stack([(gt[i].astype(np.float64) - mid_p[i]) / max_l[i] + 0.5 for i in range(n)])
    got = jn.normalize(gt, max_l, mid_p)
    assert got.shape == (n, 21, 3) and got.dtype == np.float64
    assert np.abs(got - want).max() <= 1e-6
    assert np.abs(jn.normalize(gt.reshape(n, 63), max_l, mid_p) - want.reshape(n, 63)).max() <= 1e-6
    # read_joint(): <subject>/joint_nor/<gesture>.npy next to ground_truth/ and TSDF/
    sub = tmp_path / "P0"
    (sub / "ground_truth").mkdir(parents=True)
    (sub / "TSDF").mkdir()
    np.save(sub / "ground_truth" / "1.npy", gt)
    np.savez(sub / "TSDF" / "1.npz", tsdf=np.zeros((n, 3, 2, 2, 2)), max_l=max_l, mid_p=mid_p)
    monkeypatch.setattr(jn, "DATA_dir", str(tmp_path))
    jn.read_joint()
    assert np.abs(np.load(sub / "joint_nor" / "1.npy") - want).max() <= 1e-6


def test_time_test_module(tmp_path, forloop_golden, capsys):
    # pre/time_test.py:13-30: both entry points on one .bin frame, the reference's two printed lines
    tt = import_pre("time_test")
    c = forloop_golden[0]
    raw = tmp_path / "000000_depth.bin"
    with open(raw, "wb") as f:
        f.write(np.asarray(c["header"], dtype=np.int32).tobytes())
        f.write(np.asarray(c["depth"], dtype=np.float32).tobytes())
    np.save(tmp_path / "pc.npy", c["pc"][None])
    (t1, ml1, mp1), (t2, ml2, mp2) = tt.read_data(str(raw), str(tmp_path / "pc.npy"))
    out = capsys.readouterr().out
    assert "for loop time:" in out and "numba time:" in out
    assert _same(t1, c["tsdf"]) and ml1 == c["max_l"]
    assert t2.shape == (3, 32, 32, 32) and ml2 > 0


# ----------------------------------------------------------------------------- device-resident path
def test_device_entry_point_matches_host_entry_point(vox):
    import torch
    frames = synth.make_frames(range(300, 340))
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    hv, hml, hmp = vox.voxelize_packed(headers, offsets, packed)
    dev = torch.device("cuda:0")
    d_vol, d_ml, d_mp = pkg.voxelize_device(torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev),
                                            torch.from_numpy(headers).to(dev), max_px)
    torch.cuda.synchronize()
    assert _same(d_vol.cpu().numpy(), hv)
    assert np.array_equal(d_ml.cpu().numpy(), hml) and np.array_equal(d_mp.cpu().numpy(), hmp)
    # the staging hint never changes results (0 = unknown, or too small -> direct path)
    for hint in (0, 16):
        v2, _, _ = pkg.voxelize_device(torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev),
                                       torch.from_numpy(headers).to(dev), hint)
        torch.cuda.synchronize()
        assert _same(v2.cpu().numpy(), hv)


def test_full_size_batch_properties():
    """BASELINE.json configs[1] size (8,500 frames): size-independent properties."""
    import torch
    n = 8500
    headers, depth = synth.make_uniform_batch(n, seed0=7)
    dev = torch.device("cuda:0")
    offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
    assert packed_len == depth.size
    d_depth = torch.from_numpy(depth.reshape(-1)).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_hdr = torch.from_numpy(headers).to(dev)
    cfg0 = pkg.capi.default_config(layout=pkg.LAYOUT_CXYZ)
    cfg1 = pkg.capi.default_config(layout=pkg.LAYOUT_CZYX)
    v0, ml0, mp0 = pkg.voxelize_device(d_depth, d_off, d_hdr, max_px, cfg0)
    v0b, _, _ = pkg.voxelize_device(d_depth, d_off, d_hdr, max_px, cfg0)
    v1, ml1, mp1 = pkg.voxelize_device(d_depth, d_off, d_hdr, max_px, cfg1)
    torch.cuda.synchronize()
    assert torch.equal(v0, v0b)                                    # deterministic / idempotent
    assert torch.equal(v0, v1.permute(0, 1, 4, 3, 2))              # the two layouts are transposes
    assert torch.equal(ml0, ml1) and torch.equal(mp0, mp1)
    assert float(v0.abs().max()) <= 1.0 and not bool(torch.isnan(v0).any())
    # a spot sample against the C oracle
    pick = np.arange(0, n, 271)
    cv, cml, _, _ = c_oracle.dtsdf_batch(headers[pick], offsets[: pick.size], depth[pick].reshape(-1))
    assert _same(v0[torch.from_numpy(pick).to(dev)].cpu().numpy(), cv)
    assert np.array_equal(ml0.cpu().numpy()[pick], cml)
    # a frame's volume does not depend on its neighbours: permuting frames permutes volumes
    perm = np.random.default_rng(0).permutation(n)
    vp, _, _ = pkg.voxelize_device(torch.from_numpy(depth[perm].reshape(-1)).to(dev), d_off, d_hdr, max_px, cfg0)
    torch.cuda.synchronize()
    assert torch.equal(vp, v0[torch.from_numpy(perm).to(dev)])


def test_full_msra_sized_set_on_one_gpu():
    """BASELINE.json configs[2] size on one GPU: 76,375 frames (30 GB of volumes, resident).  The offset
    table walks 8,500 distinct crops nine times over, so frame i must equal frame i mod 8,500 - a
    frame's volume depends on nothing but its own record - and a sample is checked against the C oracle."""
    import torch
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info(dev)
    n, base = 76375, 8500
    if free < 40 * (1 << 30):
        pytest.skip("needs 40 GB of free device memory")
    headers, depth = synth.make_uniform_batch(base, seed0=21)
    offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
    idx = np.arange(n) % base
    d_depth = torch.from_numpy(depth.reshape(-1)).to(dev)
    d_off = torch.from_numpy(offsets[idx]).to(dev)
    d_hdr = torch.from_numpy(headers[idx]).to(dev)
    cfg = pkg.capi.default_config(layout=pkg.LAYOUT_CXYZ)
    vol, ml, mp = pkg.voxelize_device(d_depth, d_off, d_hdr, max_px, cfg)
    torch.cuda.synchronize()
    assert vol.shape == (n, 3, 32, 32, 32)
    first = vol[:base]
    for k in range(1, (n + base - 1) // base):
        part = vol[k * base:(k + 1) * base]
        assert torch.equal(part, first[:part.shape[0]]), "replica %d" % k
    assert torch.equal(ml[base:2 * base], ml[:base]) and torch.equal(mp[-100:], mp[(n - 100) % base:(n - 100) % base + 100])
    pick = np.arange(0, base, 97)
    cv, cml, _, _ = c_oracle.dtsdf_batch(headers[pick], offsets[:pick.size], depth[pick].reshape(-1))
    assert _same(first[torch.from_numpy(pick).to(dev)].cpu().numpy(), cv)
    assert np.array_equal(ml.cpu().numpy()[pick + 7 * base], cml)
    del vol, first
    torch.cuda.empty_cache()


def test_normalize_joints(vox):
    # (gt - mid_p)/max_l + 0.5: pre/joint_nor.py:8-18 (float64 there, float32 here: 1e-6)
    rng = np.random.default_rng(3)
    gt = rng.uniform(-150, 150, size=(50, 63)).astype(np.float32)
    max_l = rng.uniform(100, 250, size=50).astype(np.float32)
    mid_p = rng.uniform(-50, 50, size=(50, 3)).astype(np.float32)
    out = vox.normalize_joints(gt, max_l, mid_p)
    ref = orc.normalize_joints(gt, max_l, mid_p)
    assert np.max(np.abs(out - ref)) <= 1e-6
    f32 = (gt.reshape(50, 21, 3) - mid_p[:, None, :]) / max_l[:, None, None] + np.float32(0.5)
    assert np.array_equal(out, f32.reshape(50, 63))
    clamped = vox.normalize_joints(gt, max_l, mid_p, clamp01=True)   # 3D_CNN/train.py:241-242
    assert np.array_equal(clamped, np.clip(f32.reshape(50, 63), 0, 1))


# ----------------------------------------------------------------------------- peer stores (NVLink)
def test_kernel_stores_into_peer_memory():
    """the gather to a training rank is done by the voxeliser kernel itself,
    storing into a buffer that lives in another GPU's HBM (gather.PeerVolumeBuffer does this across
    processes with CUDA IPC; here both GPUs belong to this process)."""
    import importlib
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    gmod = importlib.import_module(pkg.__name__ + ".gather")
    frames = synth.make_frames(range(400, 464))
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    gmod.enable_peer_access(1, 0)
    out0 = torch.zeros((len(frames), 3, 32, 32, 32), dtype=torch.float32, device="cuda:0")   # lives on GPU 0
    d1 = torch.device("cuda:1")
    with torch.cuda.device(d1):
        pkg.voxelize_device(torch.from_numpy(packed).to(d1), torch.from_numpy(offsets).to(d1),
                            torch.from_numpy(headers).to(d1), max_px, out_vol=out0)           # runs on GPU 1
        torch.cuda.synchronize(d1)
    cv, _, _, _ = c_oracle.dtsdf_batch(*pack_simple(frames))
    assert _same(out0.cpu().numpy(), cv)


def test_joint_normalisation_fused_into_the_voxeliser_launch(vox):
    """dtsdf_voxelize_joints_batch: the joints come out normalised with the frame's own mid_p / max_l
    (pre/joint_nor.py:8-18), bit-identical to the stand-alone kernel, with the trainer's clamp
    (3D_CNN/train.py:241-242) and the loader's z flip (3D_CNN/dataset.py:107) as flags."""
    import torch
    frames = synth.make_frames(range(600, 640))
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    dev = torch.device("cuda:0")
    gt = np.random.default_rng(5).uniform(-150, 150, size=(len(frames), 63)).astype(np.float32)
    args = (torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev), max_px)
    vol0, ml0, mp0 = pkg.voxelize_device(*args)
    for flags in (0, pkg.capi.JOINT_CLAMP01, pkg.capi.JOINT_FLIP_Z, pkg.capi.JOINT_CLAMP01 | pkg.capi.JOINT_FLIP_Z):
        vol, ml, mp, jn = pkg.voxelize_device(*args, joints=torch.from_numpy(gt).to(dev), joint_flags=flags)
        torch.cuda.synchronize()
        assert torch.equal(vol, vol0) and torch.equal(ml, ml0) and torch.equal(mp, mp0)
        g = gt.reshape(-1, 21, 3).copy()
        if flags & pkg.capi.JOINT_FLIP_Z:
            g[:, :, 2] = -g[:, :, 2]
        ref = vox.normalize_joints(g.reshape(-1, 63), ml.cpu().numpy(), mp.cpu().numpy(),
                                   clamp01=bool(flags & pkg.capi.JOINT_CLAMP01))
        assert np.array_equal(jn.cpu().numpy(), ref)


def test_on_the_fly_loader_overlapped_batches(vox):
    """loader.OnTheFlyLoader (BASELINE.json configs[4]): batches voxelised on a side stream from frames
    resident in HBM, joint normalisation fused; every batch equals the direct result, in order and shuffled."""
    import importlib
    import torch
    lmod = importlib.import_module(pkg.__name__ + ".loader")
    frames = synth.make_frames(range(700, 700 + 70))
    gt = np.random.default_rng(9).uniform(-150, 150, size=(len(frames), 63)).astype(np.float32)
    res = lmod.ResidentFrames.from_frames(frames, joints=gt)
    want_vol, want_ml, want_mp = vox.voxelize_frames(frames)
    want_jn = vox.normalize_joints(gt, want_ml, want_mp, clamp01=True)
    for shuffle in (False, True):
        loader = lmod.OnTheFlyLoader(res, batch_size=16, shuffle=shuffle, seed=3)
        assert len(loader) == 5
        order = loader._order().cpu().numpy()
        seen = 0
        acc = torch.zeros((), device="cuda")
        for tsdf, jn, max_l, mid_p in loader:
            m = tsdf.shape[0]
            idx = order[seen:seen + m]
            acc += tsdf.abs().sum()                   # consumer work on the current stream
            assert _same(tsdf.cpu().numpy(), want_vol[idx])
            assert np.array_equal(jn.cpu().numpy(), want_jn[idx])
            assert np.array_equal(max_l.cpu().numpy(), want_ml[idx]) and np.array_equal(mid_p.cpu().numpy(), want_mp[idx])
            seen += m
        assert seen == len(frames) and float(acc) > 0


def test_on_the_fly_loader_consumer_stream_and_previous_batch(vox):
    """The loader's contract: a batch stays valid during the iteration that received it and the next one,
    ordered on whatever stream is current in the consumer - here a stream of its own, with the previous
    batch used (delayed) together with the current one and no host synchronisation in the loop."""
    import importlib
    import torch
    lmod = importlib.import_module(pkg.__name__ + ".loader")
    frames = synth.make_frames(range(1700, 1700 + 96))
    res = lmod.ResidentFrames.from_frames(frames)
    want_vol = torch.from_numpy(vox.voxelize_frames(frames)[0]).cuda()
    loader = lmod.OnTheFlyLoader(res, batch_size=8)
    mine = torch.cuda.Stream()
    spin = torch.empty(1 << 24, device="cuda")
    sums, prev = [], None
    with torch.cuda.stream(mine):
        for tsdf, jn, max_l, mid_p in loader:
            for _ in range(4):
                spin.add_(1.0)                            # keeps the consumer's stream behind the side stream
            cur = tsdf.double().abs().sum()
            sums.append(cur if prev is None else cur + prev.double().abs().sum())   # previous batch, one iteration late
            prev = tsdf
    mine.synchronize()
    ref = [want_vol[8 * k:8 * k + 8].double().abs().sum() for k in range(12)]
    for k in range(12):
        w = ref[k] if k == 0 else ref[k] + ref[k - 1]
        assert float(sums[k]) == float(w), k


# ----------------------------------------------------------------------------- augmentation (N2)
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("R", [32, 48])
def test_augmented_dtsdf_matches_its_specification(R, layout):
    """dtsdf_voxelize_aug_batch vs oracle/dtsdf_oracle.c:oracle_dtsdf_aug_frame (the specification;
    parity with the reference is unpinned: its data_aug raises, pre/process.py:252), bit for bit, on
    random stretch + rotation about the origin and about a point inside the hand."""
    import importlib
    import torch
    amod = importlib.import_module(pkg.__name__ + ".augment")
    frames = synth.make_frames(range(800, 824)) + [synth.make_frame(1, box=(0, 0, 320, 240))]   # last: unstaged
    headers, offsets, packed = pack_simple(frames)
    n = len(frames)
    centers = np.zeros((n, 3), np.float32)
    centers[n // 2:] = [5.0, -12.0, -370.0]
    aug = amod.make_augmentations(n, np.random.default_rng(R + layout), centers=centers)
    dev = torch.device("cuda:0")
    cfg = pkg.capi.default_config(R, layout, pkg.BBOX_PIXELS)
    vol, ml, mp = amod.voxelize_aug_device(torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev),
                                           torch.from_numpy(headers).to(dev), torch.from_numpy(aug).to(dev), 130 * 150, cfg)
    torch.cuda.synchronize()
    cv, cml, cmp_ = c_oracle.dtsdf_aug_batch(headers, offsets, packed, aug, R=R, layout=layout)
    assert np.array_equal(ml.cpu().numpy(), cml) and np.array_equal(mp.cpu().numpy(), cmp_)
    got = vol.cpu().numpy()
    for i in range(n):
        assert _same(got[i], cv[i]), "frame %d" % i
    assert 0.02 < float(np.mean((got != 0) & (np.abs(got) < 1))) < 0.5       # a real surface band is in there


def test_augmented_dtsdf_unusual_transforms():
    """The augmented kernel's far-surface shortcut must never change a result: transforms it was not
    meant for - strong anisotropic stretch, a nearly singular A, an Ainv that is not the inverse of A,
    a centre far from the hand, huge / tiny scales, NaN entries - and crops with huge, negative and
    NaN depths, all bit for bit against the specification."""
    import importlib
    import torch
    amod = importlib.import_module(pkg.__name__ + ".augment")
    rng = np.random.default_rng(77)
    base = synth.make_frames(range(1200, 1212))
    mats = []

    def rec(A, Ainv=None, c=(0, 0, 0)):
        A = np.asarray(A, dtype=np.float64).reshape(3, 3)
        Ainv = np.linalg.inv(A) if Ainv is None else np.asarray(Ainv, dtype=np.float64).reshape(3, 3)
        return np.concatenate([A.reshape(-1), Ainv.reshape(-1), np.asarray(c, dtype=np.float64)]).astype(np.float32)

    def rot(ax, deg):
        t = np.deg2rad(deg)
        cs, sn = np.cos(t), np.sin(t)
        i, j = [(1, 2), (0, 2), (0, 1)][ax]
        M = np.eye(3)
        M[i, i], M[i, j], M[j, i], M[j, j] = cs, -sn, sn, cs
        return M

    mats.append(rec(np.diag([3.0, 0.3, 1.0]) @ rot(2, 50) @ rot(0, 35)))                 # strong stretch
    mats.append(rec(np.diag([1.0, 1.0, 1e-3]) @ rot(1, 20)))                             # nearly singular
    mats.append(rec(np.diag([1e-4, 1.0, 1.0])))                                          # nearly singular, axis-aligned
    mats.append(rec(rot(2, 30), Ainv=rot(2, -25)))                                       # Ainv is not A^-1
    mats.append(rec(np.diag([1.2, 0.8, 1.0]), Ainv=np.eye(3) * 1.5))                     # Ainv unrelated to A
    mats.append(rec(rot(0, 20) @ rot(1, -15), c=(4000.0, -9000.0, 20000.0)))             # centre far away
    mats.append(rec(rot(2, 10), c=(3e6, 0, -2e6)))                                       # centre beyond the guard
    mats.append(rec(np.eye(3) * 900.0))                                                  # huge scale
    mats.append(rec(np.eye(3) * 2e-3))                                                   # tiny scale
    nanA = rec(rot(2, 15)); nanA[4] = np.nan
    mats.append(nanA)                                                                    # NaN entry
    mats.append(rec(np.diag([-1.0, 1.0, 1.0]) @ rot(2, 33), c=(5, -12, -370)))           # reflection
    mats.append(rec(rot(0, 90) @ rot(2, 45), c=(5, -12, -370)))                          # depth axis swapped
    frames = []
    for i, f in enumerate(base):
        d = f["depth"].copy()
        if i % 4 == 1:
            d[rng.integers(0, d.size, 40)] = 3e12                                        # beyond 2^40
            d[rng.integers(0, d.size, 40)] = -500.0
        if i % 4 == 2:
            d[rng.integers(0, d.size, 25)] = np.nan
        frames.append({"header": f["header"], "depth": d})
    headers, offsets, packed = pack_simple(frames)
    aug = np.stack(mats).astype(np.float32)
    assert aug.shape == (len(frames), 21)
    dev = torch.device("cuda:0")
    for layout in (0, 1):
        cfg = pkg.capi.default_config(32, layout, pkg.BBOX_PIXELS)
        vol, ml, mp = amod.voxelize_aug_device(torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev),
                                               torch.from_numpy(headers).to(dev), torch.from_numpy(aug).to(dev),
                                               130 * 150, cfg)
        torch.cuda.synchronize()
        cv, cml, cmp_ = c_oracle.dtsdf_aug_batch(headers, offsets, packed, aug, R=32, layout=layout)
        assert _same(ml.cpu().numpy(), cml) and _same(mp.cpu().numpy(), cmp_)
        got = vol.cpu().numpy()
        for i in range(len(frames)):
            assert _same(got[i], cv[i]), "frame %d layout %d" % (i, layout)


@pytest.mark.gpu
@pytest.mark.parametrize("focal", [241.42, 300.0, 1e-7])
def test_augmented_dtsdf_either_side_of_the_safe_frame_test(focal):
    """The augmented kernel runs a frame through its straight-line reciprocal division only when the whole grid has
    2^-60 <= |v_z| <= 2^60 and the focal is inside 2^+-20 (csrc/dtsdf_kernels.cu: aug_grid_is_safe); both
    instantiations must give the specification's result.  Frames on either side of that test: ordinary maps (safe),
    an Ainv that stretches the grid across the camera plane z = 0 (v_z of either sign, zero and next to zero:
    infinite and NaN projections), a grid wholly behind the camera, a crop that starts at pixel (0, 0) (where the
    pixel of a NaN projection would land) - with the default focal, another one, and one outside the proven range."""
    import importlib
    import torch
    amod = importlib.import_module(pkg.__name__ + ".augment")

    def rec(A, Ainv=None, c=(0, 0, -380.0)):
        A = np.asarray(A, dtype=np.float64).reshape(3, 3)
        Ainv = np.linalg.inv(A) if Ainv is None else np.asarray(Ainv, dtype=np.float64).reshape(3, 3)
        return np.concatenate([A.reshape(-1), Ainv.reshape(-1), np.asarray(c, dtype=np.float64)]).astype(np.float32)

    t = np.deg2rad(25.0)
    Rz = np.array([[np.cos(t), np.sin(t), 0], [-np.sin(t), np.cos(t), 0], [0, 0, 1.0]])
    mats = [rec(np.diag([1.2, 0.8, 1.0]) @ Rz),                                  # safe
            rec(np.eye(3), Ainv=np.diag([1.0, 1.0, 6.0])),                       # the grid crosses z = 0
            rec(Rz, Ainv=np.diag([1.0, 1.0, 3.5]) @ Rz.T),                       # crosses z = 0, rotated
            rec(np.eye(3), Ainv=np.diag([1.0, 1.0, -1.0])),                      # grid mirrored to behind the camera (z > 0): safe
            rec(np.eye(3), Ainv=np.zeros((3, 3)), c=(0, 0, 0)),                  # every voxel at the camera centre: v_z = 0
            rec(np.diag([0.9, 1.1, 1.0]) @ Rz, c=(3.0, -2.0, -400.0))]           # safe
    frames = synth.make_frames(range(1300, 1300 + len(mats)), box=(0, 0, 110, 120))   # crops that start at pixel (0, 0)
    frames += synth.make_frames(range(1400, 1400 + len(mats)))
    aug = np.stack(mats + mats).astype(np.float32)
    headers, offsets, packed = pack_simple(frames)
    dev = torch.device("cuda:0")
    for layout in (0, 1):
        cfg = pkg.capi.default_config(32, layout, pkg.BBOX_PIXELS)
        cfg.focal = focal
        vol, ml, mp = amod.voxelize_aug_device(torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev),
                                               torch.from_numpy(headers).to(dev), torch.from_numpy(aug).to(dev),
                                               130 * 150, cfg)
        torch.cuda.synchronize()
        cv, cml, cmp_ = c_oracle.dtsdf_aug_batch(headers, offsets, packed, aug, R=32, layout=layout, focal=focal)
        assert _same(ml.cpu().numpy(), cml) and _same(mp.cpu().numpy(), cmp_)
        got = vol.cpu().numpy()
        for i in range(len(frames)):
            assert _same(got[i], cv[i]), "frame %d layout %d" % (i, layout)
        assert focal < 1.0 or (np.count_nonzero(cv[0]) > 0 and np.count_nonzero(cv[len(mats)]) > 0)


def test_identity_augmentation_is_the_plain_path(vox):
    import importlib
    import torch
    amod = importlib.import_module(pkg.__name__ + ".augment")
    frames = synth.make_frames(range(900, 916))
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    dev = torch.device("cuda:0")
    gt = np.random.default_rng(2).uniform(-150, 150, size=(len(frames), 63)).astype(np.float32)
    args = (torch.from_numpy(packed).to(dev), torch.from_numpy(offsets).to(dev), torch.from_numpy(headers).to(dev))
    vol, ml, mp, jn = amod.voxelize_aug_device(*args, torch.from_numpy(amod.identity_augmentations(len(frames))).to(dev),
                                               max_px, joints=torch.from_numpy(gt).to(dev),
                                               joint_flags=pkg.capi.JOINT_CLAMP01)
    want = pkg.voxelize_device(*args, max_px, joints=torch.from_numpy(gt).to(dev), joint_flags=pkg.capi.JOINT_CLAMP01)
    torch.cuda.synchronize()
    assert _same(vol.cpu().numpy(), want[0].cpu().numpy())
    assert torch.equal(ml, want[1]) and torch.equal(mp, want[2]) and torch.equal(jn, want[3])


def test_on_the_fly_loader_with_augmentation():
    """BASELINE.json configs[4]: on-the-fly D-TSDF + 3-D augmentation on a side stream; every batch equals
    the augmentation specification for the matrices the loader drew."""
    import importlib
    lmod = importlib.import_module(pkg.__name__ + ".loader")
    frames = synth.make_frames(range(950, 950 + 20))
    headers, offsets, packed = pack_simple(frames)
    gt = np.random.default_rng(4).uniform(-100, 100, size=(len(frames), 63)).astype(np.float32)
    res = lmod.ResidentFrames.from_frames(frames, joints=gt)
    loader = lmod.OnTheFlyLoader(res, batch_size=8, shuffle=False, augment=True, augment_center=(0.0, 0.0, -380.0))
    seen = 0
    for k, (tsdf, jn, max_l, mid_p) in enumerate(loader):
        m = tsdf.shape[0]
        aug = loader.last_aug[k % loader.NBUF]
        sel = np.arange(seen, seen + m)
        sub = [frames[i] for i in sel]
        h, o, p = pack_simple(sub)
        cv, cml, cmp_ = c_oracle.dtsdf_aug_batch(h, o, p, aug)
        assert _same(tsdf.cpu().numpy(), cv) and np.array_equal(max_l.cpu().numpy(), cml)
        assert jn.shape == (m, 63) and float(jn.min()) >= 0.0 and float(jn.max()) <= 1.0
        seen += m
    assert seen == len(frames)


# ----------------------------------------------------------------------------- randomised stress
def _random_frames(rng, n):
    """Frames that look nothing like a hand: random boxes anywhere in (and partly outside) the image,
    depth fields mixing planes, noise, holes, values straddling the |d| = 1 validity threshold and
    negative depths."""
    frames = []
    for _ in range(n):
        bw, bh = int(rng.integers(1, 180)), int(rng.integers(1, 160))
        l, t = int(rng.integers(-20, 300)), int(rng.integers(-20, 220))
        yy, xx = np.mgrid[0:bh, 0:bw].astype(np.float32)
        kind = int(rng.integers(0, 5))
        base = np.float32(rng.uniform(150, 900))
        if kind == 0:
            d = base + rng.uniform(-2, 2) * xx + rng.uniform(-2, 2) * yy
        elif kind == 1:
            d = base + rng.normal(0, 40, size=(bh, bw))
        elif kind == 2:
            d = rng.uniform(0.0, 2.0, size=(bh, bw))                       # straddles |d| < 1
        elif kind == 3:
            d = -(base + 30 * np.sin(xx / 7.0) * np.cos(yy / 5.0))         # negative depths are valid
        else:
            d = base + 80 * ((xx // 8 + yy // 8) % 2)                      # steps
        d = d.astype(np.float32)
        d[rng.random((bh, bw)) < rng.uniform(0, 0.7)] = 0.0                # holes
        frames.append({"header": np.array([320, 240, l, t, l + bw, t + bh], np.int32), "depth": d.reshape(-1)})
    return frames


@pytest.mark.parametrize("seed", [1, 2])
def test_randomised_frames_all_bbox_modes(vox, seed):
    rng = np.random.default_rng(seed)
    frames = _random_frames(rng, 300)
    headers, offsets, packed = pack_simple(frames)
    n = len(frames)
    for layout in (0, 1):
        # bbox over valid pixels
        vol, ml, mp = vox.voxelize_packed(headers, offsets, packed, layout=layout)
        cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed, layout=layout)
        for i in range(n):
            assert _same(vol[i], cv[i]), "pixels frame %d layout %d" % (i, layout)
        assert _same(ml, cml) and _same(mp, cmp_)
        # caller-supplied extrema: random boxes around a random point, from tiny to huge
        centre = rng.uniform(-300, 300, size=(n, 3)).astype(np.float32)
        centre[:, 2] = -rng.uniform(100, 1000, size=n).astype(np.float32)
        half = (10.0 ** rng.uniform(-3, 3.5, size=(n, 3))).astype(np.float32)
        aux = np.concatenate([centre + half, centre - half], axis=1).astype(np.float32)
        vol, ml, mp = vox.voxelize_packed(headers, offsets, packed, layout=layout, bbox_mode=pkg.BBOX_EXTREMA, aux=aux)
        cv, cml, cmp_, _ = c_oracle.dtsdf_batch(headers, offsets, packed, layout=layout, bbox_mode=1, aux=aux)
        for i in range(n):
            assert _same(vol[i], cv[i]), "extrema frame %d layout %d" % (i, layout)
        assert _same(ml, cml) and _same(mp, cmp_)
        # caller-supplied grid: voxel_len from 1e-4 to 1e3 mm, trunc not tied to it
        grid = np.concatenate([centre - half, (10.0 ** rng.uniform(-4, 3, size=(n, 1))),
                               (10.0 ** rng.uniform(-3, 3, size=(n, 1)))], axis=1).astype(np.float32)
        vol, _, _ = vox.voxelize_packed(headers, offsets, packed, layout=layout, bbox_mode=pkg.BBOX_GRID, aux=grid)
        cv, _, _, _ = c_oracle.dtsdf_batch(headers, offsets, packed, layout=layout, bbox_mode=2, aux=grid)
        for i in range(n):
            assert _same(vol[i], cv[i]), "grid frame %d layout %d" % (i, layout)


def test_concurrent_host_threads_with_their_own_contexts():
    """The library is re-entrant: several host threads, each with its own context (own streams, staging
    buffers and expansion pool), voxelise different batches at the same time (ctypes drops the GIL)."""
    import threading
    jobs = []
    for t in range(4):
        frames = synth.make_frames(range(1200 + 100 * t, 1200 + 100 * t + 60 + 7 * t))
        jobs.append(pack_simple(frames))
    want = [c_oracle.dtsdf_batch(h, o, p, layout=t % 2)[0] for t, (h, o, p) in enumerate(jobs)]
    got, errs = [None] * 4, []

    def work(t):
        try:
            v = pkg.Voxelizer(device=0)
            h, o, p = jobs[t]
            for _ in range(3):
                got[t] = v.voxelize_packed(h, o, p, layout=t % 2)[0]
            v.close()
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for t in range(4):
        assert _same(got[t], want[t]), "thread %d" % t
