"""GPU tests of the lossless volume codec on the device (dtsdf_pack_volumes_batch /
dtsdf_unpack_volumes_batch) and of the packed gather built on it (gather.PackedVolumeGather)."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, pkg, synth

pytestmark = pytest.mark.gpu


def _codec():
    return importlib.import_module(pkg.__name__ + ".codec")


def _reference_pack(vol):
    """The format of include/dtsdf_b200.h, restated in NumPy: codes uint8[n, R^3/4], counts, triples."""
    n = vol.shape[0]
    bits = np.ascontiguousarray(vol).view(np.uint32).reshape(n, 3, -1)
    a, b, c = bits[:, 0], bits[:, 1], bits[:, 2]
    same = (a == b) & (b == c)
    code = np.full(a.shape, 3, dtype=np.uint8)
    code[same & (a == 0)] = 0
    code[same & (a == 0x3F800000)] = 1
    code[same & (a == 0xBF800000)] = 2
    q = code.reshape(n, -1, 4)
    codes = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)
    counts = (code == 3).sum(axis=1).astype(np.int32)
    lits = [bits[i][:, code[i] == 3].T.reshape(-1) for i in range(n)]       # (x, y, z) triples in voxel order
    return codes, counts, lits


@pytest.mark.parametrize("R", [32, 48, 64])
def test_pack_matches_the_format_and_round_trips(R):
    import torch
    codec = _codec()
    frames = synth.make_frames(range(900, 900 + (24 if R == 32 else 6)))
    vox = pkg.Voxelizer(device=0, voxel_res=R)
    vol, _, _ = vox.voxelize_frames(frames)
    d_vol = torch.from_numpy(vol).cuda()
    packed = codec.pack_volumes_device(d_vol)
    torch.cuda.synchronize()
    codes, counts, lits = _reference_pack(vol)
    assert np.array_equal(packed.codes.cpu().numpy(), codes)
    assert np.array_equal(packed.lit_count.cpu().numpy(), counts)
    got = packed.lits.cpu().numpy().view(np.uint32)
    for i in range(len(frames)):
        assert np.array_equal(got[i, :3 * counts[i]], lits[i])
    back = codec.unpack_volumes_device(packed)
    assert torch.equal(back.view(torch.int32), d_vol.view(torch.int32))
    assert packed.bytes_in_use() == len(frames) * (R ** 3 // 4 + 4) + 12 * int(counts.sum())
    vox.close()


def test_round_trip_of_adversarial_volumes():
    """Every voxel a literal (the staging buffer of the expansion kernel overflows: the direct path), no
    literal at all, -0.0 / NaN payloads / denormals, mixed channels of special values, literal runs that
    start and end on every lane."""
    import torch
    codec = _codec()
    rng = np.random.default_rng(5)
    R, n = 32, 12
    r3 = R ** 3
    vol = np.zeros((n, 3, r3), dtype=np.float32)
    vol[0] = rng.standard_normal((3, r3)).astype(np.float32)                 # all literals
    vol[1] = 0.0                                                             # all code 0
    vol[2] = 1.0
    vol[3] = -1.0
    special = np.array([0.0, -0.0, 1.0, -1.0, np.nan, np.inf, 1e-42, 0.5], dtype=np.float32)
    vol[4] = special[rng.integers(0, special.size, size=(3, r3))]            # mixed channels of special values
    vol[5] = special[rng.integers(0, 4, size=(1, r3))].repeat(3, axis=0)     # same value in all channels (-0.0: literal)
    for i in range(6, n):                                                    # literal density sweep incl. runs
        dens = [0.001, 0.05, 0.2, 0.3, 0.6, 0.97][i - 6]
        base = np.array([0.0, 1.0, -1.0], dtype=np.float32)[rng.integers(0, 3, size=r3)]
        vol[i] = base[None].repeat(3, axis=0)
        m = rng.random(r3) < dens
        vol[i][:, m] = rng.standard_normal((3, int(m.sum()))).astype(np.float32)
    nanbits = vol.view(np.uint32)
    nanbits[4, 0, :7] = [0x7fc00001, 0xffc12345, 0x7f800001, 0x80000000, 0x00000001, 0x807fffff, 0x7f7fffff]
    vol = vol.reshape(n, 3, R, R, R)
    d_vol = torch.from_numpy(vol).cuda()
    packed = codec.pack_volumes_device(d_vol)
    codes, counts, lits = _reference_pack(vol)
    assert np.array_equal(packed.codes.cpu().numpy(), codes)
    assert np.array_equal(packed.lit_count.cpu().numpy(), counts)
    assert counts[0] == r3 and counts[1] == counts[2] == counts[3] == 0
    back = codec.unpack_volumes_device(packed)
    assert torch.equal(back.view(torch.int32), d_vol.view(torch.int32))
    # a count that contradicts the codes (corrupt input) must not fault
    packed.lit_count.zero_()
    codec.unpack_volumes_device(packed)
    torch.cuda.synchronize()


def test_full_size_round_trip_and_packed_size():
    """BASELINE.json configs[1] size: 8,500 volumes, 3.34 GB dense; the packed form in use is ~8x smaller."""
    import torch
    codec = _codec()
    n = 8500
    headers, depth = synth.make_uniform_batch(n, seed0=4321)
    offsets, packed_len, max_px = pkg.capi.pack_plan(headers)
    dev = torch.device("cuda:0")
    d_vol, _, _ = pkg.voxelize_device(torch.from_numpy(depth.reshape(-1)).to(dev), torch.from_numpy(offsets).to(dev),
                                      torch.from_numpy(headers).to(dev), max_px)
    packed = codec.pack_volumes_device(d_vol)
    back = codec.unpack_volumes_device(packed)
    assert torch.equal(back.view(torch.int32), d_vol.view(torch.int32))
    used = packed.bytes_in_use()
    assert used < d_vol.numel() * 4 / 6


def test_unpack_pulls_from_peer_memory():
    """The packed arrays live on GPU 1; the expansion kernel runs on GPU 0 and reads them over NVLink."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    codec = _codec()
    gmod = importlib.import_module(pkg.__name__ + ".gather")
    frames = synth.make_frames(range(700, 764))
    headers, offsets, packed_in, max_px = pkg.pack_frames(frames)
    d1 = torch.device("cuda:1")
    with torch.cuda.device(d1):
        vol1, _, _ = pkg.voxelize_device(torch.from_numpy(packed_in).to(d1), torch.from_numpy(offsets).to(d1),
                                         torch.from_numpy(headers).to(d1), max_px)
        pk = codec.pack_volumes_device(vol1)
        torch.cuda.synchronize(d1)
    gmod.enable_peer_access(0, 1)
    with torch.cuda.device(0):
        out0 = codec.unpack_volumes_device(pk, device="cuda:0")
        torch.cuda.synchronize(0)
    assert out0.device.index == 0
    assert torch.equal(out0.cpu().view(torch.int32), vol1.cpu().view(torch.int32))


def test_packed_gather_across_processes():
    """gather.PackedVolumeGather under torchrun with one rank per GPU (tools/gather_check.py): the root's
    buffer equals every rank's local volumes."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(ngpu, 8)),
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tools", "gather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "gather_check ok" in r.stdout
