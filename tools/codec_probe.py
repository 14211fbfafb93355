"""Times the device volume codec on one GPU: pack (one pass over dense volumes) and unpack (local memory;
over NVLink it is tools/gather_check.py / bench.py --gpus N).  GPU box."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    pkg = graft.load_package()
    codec = importlib.import_module(pkg.__name__ + ".codec")
    dev = torch.device("cuda:0")
    for R, n in ((32, 8500), (48, 2500), (64, 1000)):
        headers, depth = pkg.synth.make_uniform_batch(n, seed0=1234)
        offsets, _, max_px = pkg.capi.pack_plan(headers)
        cfg = pkg.capi.default_config(R, 0, 0)
        d_vol, _, _ = pkg.voxelize_device(torch.from_numpy(depth.reshape(-1)).to(dev), torch.from_numpy(offsets).to(dev),
                                          torch.from_numpy(headers).to(dev), max_px, cfg)
        packed = codec.pack_volumes_device(d_vol)
        out = torch.empty_like(d_vol)
        used = packed.bytes_in_use()
        dense = d_vol.numel() * 4
        t_pack = timed(lambda: codec.pack_volumes_device(d_vol, packed))
        t_unpack = timed(lambda: codec.unpack_volumes_device(packed, out=out))
        assert torch.equal(out.view(torch.int32), d_vol.view(torch.int32))
        print("R=%d n=%d dense %.2f GB packed in use %.3f GB (%.1fx; %.1f KB/frame) | pack %.3f ms (%.0f GB/s read+write) | "
              "unpack %.3f ms (%.0f GB/s read+write, %.2f M volumes/s)"
              % (R, n, dense / 1e9, used / 1e9, dense / used, used / n / 1e3, t_pack, (dense + used) / t_pack / 1e6,
                 t_unpack, (dense + used) / t_unpack / 1e6, n / t_unpack / 1e3), flush=True)


if __name__ == "__main__":
    main()
