"""Host-side logic that needs no GPU: the library loads and exports the whole C ABI, argument
checking, packing, frame sharding (incl. a world_size-2 gloo run) and the synthetic generator."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as graft
from helpers import ROOT, pack_simple, pkg, synth


def test_library_exports_every_symbol_in_header():
    graft.build_cuda()
    lib = pkg.capi.load_library()
    names = graft.header_symbols()
    assert len(names) >= 19
    for s in names:
        assert hasattr(lib, s), s
    assert set(names) == set(pkg.capi._SIGNATURES)          # the ctypes table covers the whole ABI
    assert lib.dtsdf_abi_version() == 1
    assert lib.dtsdf_error_string(-4).decode().startswith("no usable CUDA device")


def test_header_is_plain_c(tmp_path):
    """include/dtsdf_b200.h is the drop-in boundary: it must compile as C99 (and C++11) on its own -
    extern "C", plain pointers and sizes, no CUDA or torch types."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "dtsdf_b200.h"\nint main(void) { return DTSDF_OK; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                   check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no module of the package imports, loads or executes anything
    under oracle/ (comments may cite it)."""
    import ast
    pkg_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "handposeestimation-with-3d-cnns_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for name in files:
            path = os.path.join(dirpath, name)
            if name.endswith(".py"):
                tree = ast.parse(open(path).read())
                for node in ast.walk(tree):
                    if isinstance(node, ast.Import):
                        assert not any(a.name.split(".")[0] == "oracle" for a in node.names), path
                    if isinstance(node, ast.ImportFrom):
                        assert (node.module or "").split(".")[0] != "oracle", path
                    if isinstance(node, ast.Constant) and isinstance(node.value, str) and "\n" not in node.value:
                        assert "libdtsdf_oracle" not in node.value and "oracle/_" not in node.value, path
            elif name.endswith((".cu", ".cpp", ".h")):
                text = open(path).read()
                assert "#include \"../../oracle" not in text and "dtsdf_oracle.c\"" not in text, path


def test_default_config_matches_reference_constants():
    cfg = pkg.capi.default_config()
    assert (cfg.voxel_res, cfg.layout, cfg.bbox_mode) == (32, 0, 0)          # pre/process.py:88
    assert (cfg.center_x, cfg.center_y) == (160, 120)                        # pre/tsdf_numba.py:9-10
    assert cfg.focal == np.float32(241.42) and cfg.trunc_mul == 3.0          # :8, :146


def test_argument_checks_need_no_device():
    lib = pkg.capi.load_library()
    cfg = pkg.capi.default_config()
    assert lib.dtsdf_voxelize_batch(None, 1, None, None, None, None, 0, None, None, None, None) == pkg.capi.ERR_INVALID_ARG
    cfg.voxel_res = 40
    assert lib.dtsdf_voxelize_batch(ctypes.byref(cfg), 1, 8, 8, 8, None, 0, 8, None, None, None) == pkg.capi.ERR_UNSUPPORTED
    cfg.voxel_res = 32
    assert lib.dtsdf_voxelize_batch(ctypes.byref(cfg), 0, None, None, None, None, 0, None, None, None, None) == 0
    assert lib.dtsdf_voxelize_batch(ctypes.byref(cfg), -1, 8, 8, 8, None, 0, 8, None, None, None) == pkg.capi.ERR_INVALID_ARG
    cfg.bbox_mode = pkg.BBOX_EXTREMA       # aux is required
    assert lib.dtsdf_voxelize_batch(ctypes.byref(cfg), 1, 8, 8, 8, None, 0, 8, None, None, None) == pkg.capi.ERR_INVALID_ARG


def test_no_cpu_fallback_without_device():
    if pkg.capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.DtsdfError) as e:
        pkg.Voxelizer(device=0)
    assert e.value.code in (pkg.capi.ERR_NO_DEVICE, pkg.capi.ERR_CUDA)
    if graft.PKG_DIR + "/pre" not in sys.path:
        sys.path.insert(0, os.path.join(graft.PKG_DIR, "pre"))
    import tsdf_numba
    fr = synth.make_frame(0, key="data")
    with pytest.raises(pkg.DtsdfError):
        tsdf_numba.cal_tsdf_cuda(fr)


def test_pack_plan_and_pack_frames():
    frames = synth.make_frames(range(5)) + [synth.make_frame(9, box=(10, 10, 13, 11))]
    headers, offsets, packed, max_px = pkg.pack_frames(frames)
    h2, o2, p2 = pack_simple(frames)
    assert np.array_equal(headers, h2) and np.array_equal(offsets, o2) and np.array_equal(packed, p2)
    assert np.all(offsets % 4 == 0)
    assert max_px == max(f["depth"].size for f in frames)
    with pytest.raises(pkg.DtsdfError) as e:
        pkg.capi.pack_plan(np.array([[320, 240, 10, 10, 10, 20]], np.int32))
    assert e.value.code == pkg.capi.ERR_BAD_HEADER
    with pytest.raises(ValueError):
        pkg.pack_frames([{"header": frames[0]["header"], "depth": frames[0]["depth"][:-3]}])
    o, total, mx = pkg.capi.pack_plan(np.zeros((0, 6), np.int32))
    assert o.size == 0 and total == 0 and mx == 0


def test_shard_range_partitions_frames():
    for n in (0, 1, 7, 8, 8500, 76375):
        for w in (1, 2, 4, 8):
            parts = [pkg.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_packed_rebases_offsets():
    frames = synth.make_frames(range(11))
    headers, offsets, packed, _ = pkg.pack_frames(frames)
    seen = 0
    for r in range(3):
        h, o, p = pkg.shard_packed(headers, offsets, packed, r, 3)
        assert o.size == h.shape[0] and (o.size == 0 or o[0] == 0)
        for i in range(h.shape[0]):
            n = frames[seen]["depth"].size
            assert np.array_equal(p[o[i]:o[i] + n], frames[seen]["depth"])
            seen += 1
    assert seen == 11


def test_synthetic_frames_are_deterministic_and_msra_shaped():
    a, b = synth.make_frame(5), synth.make_frame(5)
    assert np.array_equal(a["header"], b["header"]) and np.array_equal(a["depth"], b["depth"])
    h = a["header"]
    assert h.dtype == np.int32 and h[0] == 320 and h[1] == 240
    assert a["depth"].dtype == np.float32 and a["depth"].size == (h[4] - h[2]) * (h[5] - h[3])
    d = synth.make_frame(0, box=synth.DEFAULT_BOX)["depth"]
    assert d.size == 15600 and 8000 < np.count_nonzero(d) < 9500
    hdr, dep = synth.make_uniform_batch(10, seed0=1)
    assert hdr.shape == (10, 6) and dep.shape == (10, 15600) and dep.dtype == np.float32


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from oracle import c_oracle
    frames = synth.make_frames(range(40, 49))
    headers, offsets, packed, _ = pkg.pack_frames(frames)
    h, o, p = pkg.shard_packed(headers, offsets, packed, rank, world)
    # the per-rank work of the sharded job, computed by the oracle here (no GPU in this test)
    vol, max_l, _, _ = c_oracle.dtsdf_batch(h, o, p, nthreads=1)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([h.shape[0]], dtype=torch.int64))
    sizes = [int(c) for c in counts]
    pad = max(sizes)
    mine = torch.zeros((pad,) + vol.shape[1:])
    mine[: vol.shape[0]] = torch.from_numpy(vol)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        full = torch.cat([g[:s] for g, s in zip(gathered, sizes)]).numpy()
        ref, _, _, _ = c_oracle.dtsdf_batch(headers, offsets, packed, nthreads=1)
        q.put(bool(np.array_equal(full, ref)) and sum(sizes) == 9)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather_gloo():
    """world_size 2 over gloo: contiguous shards, rebased offsets, all-gather of the volumes to
    every rank reproduces the single-process result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


# ------------------------------------------------------------------ the caller's np.random stream, in C
_RNG_CHECK = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import __graft_entry__ as g
capi = g.load_package().capi
rng = np.random.default_rng(7)
n = %d
# numbers of valid points as MSRA crops have them, plus the corners: 1 point, powers of two and their
# neighbours (the rejection mask changes there), fewer points than point_num
n_valid = rng.integers(1, 30000, size=n).astype(np.int32)
corner = np.array([1, 2, 3, 4, 5, 255, 256, 257, 4095, 4096, 4097, 5999, 6000, 6001, 8191, 8192, 8193, 65535, 65536,
                   65537, (1 << 20) - 1, 1 << 20], dtype=np.int32)
n_valid[:corner.size] = corner
point_num = 6000
np.random.seed(20181)
np.random.random(777)                                   # not at a block boundary of the generator
state = np.random.get_state()
bad = 0
CH = 2000
for i0 in range(0, n, CH):                              # the library, a chunk of frames per call
    np.random.set_state(state)
    want = np.empty((min(CH, n - i0), point_num), dtype=np.int32)
    for i in range(want.shape[0]):                      # pre/process.py:70-84, verbatim calls
        ps = int(n_valid[i0 + i])
        if ps < point_num:
            want[i] = np.arange(0, point_num, 1, dtype=np.int32)
            want[i, ps:] = np.random.randint(0, ps, size=point_num - ps)
        else:
            want[i] = np.random.randint(0, ps, size=point_num)
    after = np.random.get_state()
    np.random.set_state(state)
    got = capi.resample_indices(n_valid[i0:i0 + want.shape[0]], point_num)
    mine = np.random.get_state()
    bad += int(not np.array_equal(got, want))
    bad += int(not (np.array_equal(mine[1], after[1]) and mine[2] == after[2]))
    state = after
# a call too small for the two-phase form (the plain form), and a frame without points in the middle of a large
# call: np.random.randint(0, 0) raises in the reference, the stream has advanced by the frames before it
def reference_draws(counts):                             # pre/process.py:70-84, verbatim calls
    rows = np.empty((len(counts), point_num), dtype=np.int32)
    for i, ps in enumerate(int(v) for v in counts):
        if ps < point_num:
            rows[i] = np.arange(0, point_num, 1, dtype=np.int32)
            rows[i, ps:] = np.random.randint(0, ps, size=point_num - ps)
        else:
            rows[i] = np.random.randint(0, ps, size=point_num)
    return rows
np.random.set_state(state)
want = reference_draws(n_valid[100:140])
after = np.random.get_state()
np.random.set_state(state)
got = capi.resample_indices(n_valid[100:140], point_num)
mine = np.random.get_state()
bad += int(not np.array_equal(got, want))
bad += int(not (np.array_equal(mine[1], after[1]) and mine[2] == after[2]))
nv = n_valid[100:700].copy()
nv[417] = 0
np.random.set_state(state)
reference_draws(nv[:417])
after = np.random.get_state()
np.random.set_state(state)
try:
    capi.resample_indices(nv, point_num)
    bad += 1
except ValueError:
    pass
mine = np.random.get_state()
bad += int(not (np.array_equal(mine[1], after[1]) and mine[2] == after[2]))
# single calls incl. count 0 and high 1 (which consumes nothing)
np.random.seed(5)
a = [np.random.randint(0, h, size=c) for h, c in ((1, 10), (7, 0), (2, 33), (1 << 30, 5), (3, 1000))]
np.random.seed(5)
b = [capi.np_randint(h, c) for h, c in ((1, 10), (7, 0), (2, 33), (1 << 30, 5), (3, 1000))]
bad += int(not all(np.array_equal(x, y) for x, y in zip(a, b)))
bad += int(np.random.randint(0, 1 << 20) != (np.random.seed(5), [np.random.randint(0, h, size=c) for h, c in
            ((1, 10), (7, 0), (2, 33), (1 << 30, 5), (3, 1000))], np.random.randint(0, 1 << 20))[2])
try:
    capi.np_randint(0, 3)
    bad += 1
except ValueError:
    pass
print("RNG_CHECK", bad)
'''


@pytest.mark.parametrize("isa,frames", [("native", 100000), ("scalar", 6000)])
def test_resampling_indices_equal_numpy_global_stream(isa, frames):
    """dtsdf_resample_indices / dtsdf_np_randint consume the caller's legacy np.random stream exactly as
    the reference's own np.random.randint calls do (pre/process.py:75-80): equal index vectors and equal
    generator state afterwards, over 10^5 frames of random point counts (both instruction-set paths)."""
    env = dict(os.environ)
    if isa == "scalar":
        env["DTSDF_RNG_ISA"] = "scalar"
    r = subprocess.run([sys.executable, "-c", _RNG_CHECK % (ROOT, frames)], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "RNG_CHECK 0" in r.stdout, r.stdout[-500:]
